"""The oracle (oracle/crvae_oracle.py) against the golden vectors produced by running the reference
itself (tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import crvae_oracle as O
from oracle.ref_loader import reference_available
from tests.conftest import GOLDEN


def _params(g, prefix):
    prm = {k: torch.from_numpy(g[prefix + k].copy()) for k in O.PARAM_KEYS}
    prm["mask"] = torch.from_numpy(g[prefix + "mask"].copy())
    return prm


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def step():
    return np.load(os.path.join(GOLDEN, "p4_step.npz"))


@pytest.fixture(scope="module")
def traj():
    return np.load(os.path.join(GOLDEN, "p10_traj.npz"))


def test_forward_matches_reference(step):
    prm = _params(step, "init.")
    act = O.crvae_forward(prm, torch.from_numpy(step["X"]), torch.from_numpy(step["eps"]))
    assert np.array_equal(act["mu"].numpy(), step["fc_mu_out"])          # encoder path is bit-exact
    assert np.array_equal(act["log_var"].numpy(), step["fc_std_out"])
    assert _rel(act["pred"].numpy(), step["pred"]) < 2e-6
    ld = O.crvae_loss(prm, act, float(step["lam_ridge"]), float(step["beta"]))
    assert abs(float(ld["loss"]) - float(step["loss"])) < 2e-6 * float(step["loss"])
    assert abs(float(ld["kl"]) - float(step["kl"])) < 2e-6 * float(step["kl"])
    assert abs(float(ld["ridge"]) - float(step["ridge"])) < 2e-6 * float(step["ridge"])
    assert abs(float(ld["smooth"]) - float(step["smooth"])) < 2e-6 * float(step["smooth"])


def test_backward_matches_autograd_of_reference(step):
    prm = _params(step, "init.")
    act = O.crvae_forward(prm, torch.from_numpy(step["X"]), torch.from_numpy(step["eps"]))
    lam_ridge, beta = float(step["lam_ridge"]), float(step["beta"])
    ld = O.crvae_loss(prm, act, lam_ridge, beta)
    g = O.crvae_backward(prm, act, ld, lam_ridge, beta)
    for k in O.PARAM_KEYS:
        assert _rel(g[k].numpy(), step["grad." + k]) < 5e-6, k


def test_gd_prox_gc_match_reference(step):
    prm = _params(step, "init.")
    grads = {k: torch.from_numpy(step["grad." + k].copy()) for k in O.PARAM_KEYS}
    O.gd_step(prm, grads, float(step["lr"]))
    assert np.array_equal(prm["w_ih"].numpy(), step["pre_prox_w_ih"])       # same op sequence -> bit-exact
    prm["w_ih"] = O.prox_update(prm["w_ih"], float(step["lam"]), float(step["lr"]))
    for k in O.PARAM_KEYS:
        assert np.array_equal(prm[k].numpy(), step["post." + k]), k
    assert np.array_equal(O.gc_matrix(prm["w_ih"]).numpy(), step["GC"])
    assert np.array_equal(O.gc_matrix(prm["w_ih"], False).numpy(), step["GC_norm"])


def test_prox_adversarial_columns():
    """Columns just above / below / at the threshold and all-zero columns."""
    lam, lr = 0.1, 5e-2
    thr = lam * lr
    g = torch.Generator().manual_seed(1)
    w = torch.randn(3, 192, 8, generator=g)
    w = w / torch.norm(w, dim=1, keepdim=True)
    scale = torch.tensor([thr * (1 + 1e-3), thr * (1 - 1e-3), thr * 0.5, thr * 2, thr * (1 + 1e-5), thr * (1 - 1e-5), 0.0, 1.0])
    w = w * scale
    out = O.prox_update(w, lam, lr)
    nz = (torch.norm(out, dim=1) > 0).numpy()
    assert nz.tolist() == [[True, False, False, True, True, False, False, True]] * 3
    assert torch.equal(out[:, :, 6], torch.zeros(3, 192))                  # 0/(lam*lr)*0 stays 0
    assert np.array_equal(O.gc_matrix(out).numpy(), nz.astype(np.int32))


def test_adam_matches_torch_optim():
    torch.manual_seed(3)
    p0 = torch.randn(257)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    prm, state = {"x": p0.clone()}, {}
    for step in range(1, 6):
        g = torch.randn(257)
        ref.grad = g.clone()
        opt.step()
        O.adam_step(prm, {"x": g}, state, step)
        assert torch.allclose(prm["x"], ref.detach(), rtol=0, atol=1e-7)


def test_free_running_trajectory_tracks_reference(traj):
    """Oracle train_phase1 (explicit backward) vs the reference's autograd run: same noise stream,
    same batch; first 101 iterations."""
    Xt = torch.from_numpy(traj["data"].T.copy())[None]
    prm = _params(traj, "init.")
    torch.manual_seed(0); np.random.seed(0)
    # consume exactly the init draws the reference model construction made
    torch.nn.GRU(10, 64); torch.nn.Linear(64, 64); torch.nn.Linear(64, 64)
    for _ in range(10):
        torch.nn.GRU(10, 64); torch.nn.Linear(64, 1)
    log = []
    O.train_phase1(prm, Xt, 20, 5e-2, 101, lam=0.1, check_every=50, log=log)
    for i, row in enumerate(log):
        assert row["it"] == int(traj["log_it"][i])
        assert abs(row["mean_loss"] - traj["log_loss"][i]) < 2e-6      # reference prints 6 decimals
        assert abs(row["kl"] - traj["log_kl"][i]) < 2e-6
        assert row["usage"] == traj["log_usage"][i]


def test_resume_from_checkpoint_reaches_next_checkpoint(traj):
    """P3: from the golden checkpoint at it=150 (prox actively zeroing columns between 150 and 250),
    50 free-running oracle iterations land on the golden checkpoint at it=200."""
    Xt = torch.from_numpy(traj["data"].T.copy())[None]
    prm = _params(traj, "ckpt150.")
    torch.set_rng_state(torch.from_numpy(traj["ckpt150._torch_rng"].copy()))
    log = []
    O.train_phase1(prm, Xt, 20, 5e-2, 201, lam=0.1, check_every=50, log=log, resume_it=150, idx=traj["idx"])
    assert [r["it"] for r in log] == [150, 200]
    assert log[1]["usage"] == traj["log_usage"][4]
    assert np.array_equal(O.gc_matrix(prm["w_ih"]).numpy().astype(np.int8), traj["log_gc"][4])


def test_golden_final_gc_hash(traj):
    gc = traj["final_GC"].astype(np.int32)
    assert hashlib.sha256(gc.tobytes()).hexdigest() == str(traj["final_GC_sha256"])
    # the survey's independent run of the same configuration
    assert str(traj["final_GC_sha256"]) == "d11a29d6dfb68f99f89c52bf4d78348b41141934ac28c9cdc7037334d1457cab"
    assert int(traj["best_it"]) == 4750
    assert np.array_equal(O.gc_matrix(torch.from_numpy(traj["final.w_ih"])).numpy(), gc)


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference_ragged():
    """Ragged (pruned) heads, as built for phase 2 (:788-790): forward + all gradients."""
    from oracle.ref_loader import load_reference
    ref = load_reference()
    p, B = 6, 32
    rng = np.random.RandomState(0)
    conn = (rng.rand(p, p) < 0.5).astype(int)
    np.fill_diagonal(conn, 1)
    torch.manual_seed(5)
    m = ref.CRVAE(p, conn, 64)
    X = torch.randn(B, 20, p)
    st = torch.get_rng_state()
    eps = torch.randn(size=(1, B, 64))[0]
    torch.set_rng_state(st)
    pred, lv, mu = m(X)
    loss = sum(torch.nn.functional.mse_loss(pred[i][:, :, 0], X[:, 10:, i]) for i in range(p))
    kl = (-0.5 * (1 + mu - lv ** 2 - torch.exp(mu)).sum(-1).sum(0)).mean(0)   # swapped names, as :482/:486
    (loss + 0.1 * kl).backward()
    prm = O.params_from_state_dict(m.state_dict(), conn)
    act = O.crvae_forward(prm, X, eps)
    ld = O.crvae_loss(prm, act, 0.0, 0.1)
    g = O.crvae_backward(prm, act, ld, 0.0, 0.1)
    gref = O.params_from_state_dict({k: v.grad for k, v in m.named_parameters()}, conn)
    assert abs(float(ld["smooth"]) - float(loss + 0.1 * kl)) < 1e-5
    for k in O.PARAM_KEYS:
        assert _rel(g[k].numpy(), gref[k].numpy()) < 5e-6, k
    sd = O.state_dict_from_params(prm)
    for k, v in m.state_dict().items():
        assert torch.equal(sd[k], v), k


def test_ref_port(step):
    """oracle/ref_port.py (the CPU-baseline port that keeps the reference's per-head structure)
    reproduces the reference's golden iteration bit-for-bit where the op sequence is the same."""
    from oracle import ref_port as RP
    p = 4
    torch.manual_seed(0); np.random.seed(0)
    m = RP.PortCRVAE(p, np.ones((p, p)), 64)
    X = torch.from_numpy(step["X"])
    torch.manual_seed(99)
    st = torch.get_rng_state()
    eps = torch.randn(size=(1, X.shape[0], 64))
    assert np.array_equal(O.params_from_state_dict(m.state_dict(), np.ones((p, p)))["w_ih"].numpy(), step["init.w_ih"])
    # feed the golden eps by rewinding: draw position differs from the golden run, so compare through the oracle
    torch.set_rng_state(st)
    smooth, loss, mmd = RP.smooth_loss(m, X, float(step["lam_ridge"]), float(step["beta"]))
    prm = _params(step, "init.")
    act = O.crvae_forward(prm, X, eps[0])
    ld = O.crvae_loss(prm, act, float(step["lam_ridge"]), float(step["beta"]))
    assert abs(float(smooth) - float(ld["smooth"])) < 2e-6 * abs(float(ld["smooth"]))
    RP.iteration(m, X, smooth, float(step["lr"]), float(step["lam"]), float(step["lam_ridge"]), float(step["beta"]))
    grads = O.crvae_backward(prm, act, ld, float(step["lam_ridge"]), float(step["beta"]))
    O.gd_step(prm, grads, float(step["lr"]))
    prm["w_ih"] = O.prox_update(prm["w_ih"], float(step["lam"]), float(step["lr"]))
    post = O.params_from_state_dict(m.state_dict(), np.ones((p, p)))
    for k in O.PARAM_KEYS:
        assert _rel(post[k].numpy(), prm[k].numpy()) < 5e-6, k


@pytest.fixture(scope="module")
def ph2():
    return np.load(os.path.join(GOLDEN, "p10_phase2.npz"))


def test_phase2_iteration_matches_reference(ph2):
    """Pruned CRVAE (golden GC as connection; ragged heads) + VRAE4E + Adam: one full phase-2 iteration."""
    conn = ph2["connection"]
    prm = {k: torch.from_numpy(ph2["c_init." + k].copy()) for k in O.PARAM_KEYS}
    prm["mask"] = torch.from_numpy(ph2["c_init.mask"].copy())
    assert np.array_equal(prm["mask"].numpy(), O.connection_mask(conn))
    vprm = {k: torch.from_numpy(ph2["v_init." + k].copy()) for k in O.VRAE_KEYS}
    wins = O.arrange_input(torch.from_numpy(np.load(os.path.join(GOLDEN, "p10_traj.npz"))["data"].T.copy()), 20)[0]
    X = wins[ph2["idx"]]
    state = {}
    r = O.phase2_iteration(prm, vprm, state, 1, X, torch.from_numpy(ph2["eps_c"]), torch.from_numpy(ph2["eps_e"]), 5e-2)
    assert abs(float(r["lossd"]["loss"]) - float(ph2["loss"])) < 2e-6 * float(ph2["loss"])
    assert abs(float(r["lossd"]["kl"]) - float(ph2["kl"])) < 2e-6 * float(ph2["kl"])
    assert _rel(r["err"].numpy(), ph2["error"]) < 2e-6
    assert _rel(r["vact"]["pred"].permute(1, 0, 2).numpy(), ph2["pred_e"]) < 5e-6
    assert abs(float(r["vloss"]["loss"]) - float(ph2["loss_e"])) < 2e-6 * float(ph2["loss_e"])
    assert abs(float(r["vloss"]["kl"]) - float(ph2["kl_e"])) < 2e-6 * float(ph2["kl_e"])
    for k in O.VRAE_KEYS:
        assert _rel(r["vgrads"][k].numpy(), ph2["v_grad." + k]) < 1e-5, k
        assert _rel(vprm[k].numpy(), ph2["v_post." + k]) < 1e-6, k           # Adam step
    for k in O.PARAM_KEYS:
        assert _rel(r["grads"][k].numpy(), ph2["c_grad." + k]) < 1e-5, k
        assert _rel(prm[k].numpy(), ph2["c_post." + k]) < 1e-6, k


def test_cs_divergence_matches_reference():
    """Oracle restatement of gaussian_overlap / cs_divergence_gmm (CR-CS-RAE.py:124-163) and the gradients of
    the trainer's lambda_cs * mean(D_CS) against the reference's own evaluation (tests/golden/cs_p10.npz)."""
    g = np.load(os.path.join(GOLDEN, "cs_p10.npz"))
    lat, pm, pl = (torch.from_numpy(g[k].copy()) for k in ("cs_lat", "cs_pm", "cs_pl"))
    vals = O.cs_divergence_gmm(lat[:, 64:], torch.exp(lat[:, :64]), pm, pl.exp())
    assert np.array_equal(vals.numpy(), g["cs_vals"])
    cs, dl, dm, dv = O.cs_head(lat, pm, pl, 0.1)
    assert abs(float(cs) - float(g["cs_vals"].mean())) < 1e-6
    assert np.array_equal(dl.numpy(), g["cs_dlat"]) and np.array_equal(dm.numpy(), g["cs_dpm"]) and np.array_equal(dv.numpy(), g["cs_dpl"])


def test_generic_vrae_matches_reference():
    """Config 4 (VRAE.py): oracle forward / loss / hand-derived backward vs the reference's autograd."""
    g = np.load(os.path.join(GOLDEN, "vrae_generic.npz"))
    prm = {k: torch.from_numpy(g["init." + k].copy()) for k in O.GVRAE_KEYS}
    a = O.gvrae_forward(prm, torch.from_numpy(g["data"]), torch.from_numpy(g["eps"]), "tanh")
    assert _rel(a["mu"].numpy(), g["mu"]) < 2e-6 and _rel(a["logvar"].numpy(), g["logvar"]) < 2e-6
    assert _rel(a["recon"].permute(1, 0, 2).numpy(), g["recon"]) < 5e-6
    l = O.gvrae_loss(a, 0.5)
    assert abs(float(l["total"]) - float(g["total"])) < 2e-6 * float(g["total"])
    assert abs(float(l["rec"]) - float(g["rec"])) < 2e-6 * float(g["rec"]) and abs(float(l["kld"]) - float(g["kld"])) < 2e-6 * float(g["kld"])
    gr = O.gvrae_backward(prm, a, l, 0.5, "tanh")
    for k in O.GVRAE_KEYS:
        assert _rel(gr[k].numpy(), g["grad." + k]) < 1e-5, k
    assert not bool(g["start_token_has_grad"])          # teacher forcing 1.0: the start token is never used (:79-82)

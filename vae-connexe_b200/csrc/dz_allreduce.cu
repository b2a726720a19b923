// One-shot all-reduce of dz = sum_heads dh0 across the head shards, FUSED with the local head sum before it and the
// latent backward after it -- one kernel over NVLink peer memory instead of  [sum kernel -> NCCL all-reduce -> pointwise kernel].
//
// The message is tiny (B x H fp32 = 64 KB, constant in p) and sits in the middle of the backward's critical
// path (decoder BPTT -> dz -> encoder BPTT), so what matters is latency, not bandwidth: every rank writes its partial into
// its own slot of a SYMMETRIC buffer (same virtual layout on every GPU, peers mapped over NVLink / NVSwitch), raises a flag
// in every peer's copy, waits for the peers' flags and reads the peers' partials directly (P2P loads), summing them in rank
// order -- the same order on every rank, so all ranks obtain bit-identical dz (the replicated encoder must not drift).
//   * chunked: CTA c owns 1024 elements end to end (local sum -> publish -> wait -> gather -> epilogue); chunks never wait
//     for each other, so no grid-wide barrier is needed;
//   * two alternating data slots + monotonically increasing epochs in the flags make the kernel re-launchable (CUDA-graph
//     replay) without a second barrier: a rank can only overwrite slot (e mod 2) at epoch e+2 after every peer signalled
//     epoch e+1, which each does after its reads of epoch e;
//   * bounded spins: a protocol failure traps instead of hanging the GPU.
// Symmetric buffer layout (32-bit words): slot0[n] | slot1[n] | flags[world][nchunks] | epoch[nchunks],  n = B*Z.
//
// Reference semantics replaced: autograd's accumulation of every head's gradient into z (CRVAE_lorenz96.py:218, :497).
#include "common.cuh"

namespace crvae {

constexpr int DZ_THREADS = 256;
constexpr int DZ_CHUNK = DZ_THREADS * 4;
constexpr int DZ_MAX_WORLD = 16;

struct DzArgs {
    const float* dh0; int P;
    float* peers[DZ_MAX_WORLD];
    int rank, world;
    const float* lat; const float* eps; float beta; int kl_form;
    float* dlat; float* dz_out;
    int B, Z, n, nchunks;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(DZ_THREADS) dz_allreduce_latent_bwd_kernel(DzArgs a) {
    const int c = blockIdx.x, tid = threadIdx.x;
    const int e0 = c * DZ_CHUNK + tid * 4;                 // this thread's 4 consecutive elements
    const bool live = e0 < a.n;                            // n is a multiple of 4 (checked by the host)
    float* mine = a.peers[a.rank];
    unsigned* my_flags = reinterpret_cast<unsigned*>(mine + 2 * (long long)a.n);
    unsigned* my_epoch = my_flags + a.world * a.nchunks;
    const unsigned e = my_epoch[c] + 1u;
    const long long slot_off = (long long)(e & 1u) * a.n;

    // 1. local sum over this rank's heads, fixed order
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live)
        for (int i0 = 0; i0 < a.P; i0 += 8) {                 // 8 loads in flight, then the adds in head order
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[j] = i0 + j < a.P ? __ldcs(reinterpret_cast<const float4*>(a.dh0 + (long long)(i0 + j) * a.n + e0)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (i0 + j < a.P) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
        }
    if (live) *reinterpret_cast<float4*>(mine + slot_off + e0) = acc;
    __threadfence_system();
    __syncthreads();
    // 2. publish: flag (epoch) into every peer's copy; 3. wait for every peer's flag in my copy
    if (tid < a.world) {
        unsigned* peer_flags = reinterpret_cast<unsigned*>(a.peers[tid] + 2 * (long long)a.n);
        st_release_sys(peer_flags + a.rank * a.nchunks + c, e);
        const unsigned* f = my_flags + tid * a.nchunks + c;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(f) - e) < 0) {
            if (++spins > (1u << 27)) __trap();
        }
    }
    __syncthreads();
    // 4. gather the partials over NVLink, summed in rank order (identical on every rank)
    // (all peer loads are issued before the first add: eight NVLink round trips in flight instead of one after the other)
    float4 dz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        // (unconditional loads -- ranks past `world` re-read this rank's own slot -- so that nothing but the loads stands between them)
#pragma unroll
        for (int half = 0; half < DZ_MAX_WORLD; half += 8) {
            if (half >= a.world) break;
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = half + j;
                v[j] = ld_volatile_f4((r < a.world ? a.peers[r] : mine) + slot_off + e0);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool on = half + j < a.world;
                dz.x += on ? v[j].x : 0.f; dz.y += on ? v[j].y : 0.f; dz.z += on ? v[j].z : 0.f; dz.w += on ? v[j].w : 0.f;
            }
        }
    }
    if (tid == 0) my_epoch[c] = e;
    if (!live) return;
    if (a.dz_out) *reinterpret_cast<float4*>(a.dz_out + e0) = dz;
    if (!a.dlat) return;
    // 5. gradient into [mu | log_var] (same arithmetic as latent_bwd_kernel, pointwise.cu)
    const float dzv[4] = {dz.x, dz.y, dz.z, dz.w};
    const float invB = 1.f / (float)a.B;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int el = e0 + j, b = el / a.Z, h = el - b * a.Z;
        const float mu = a.lat[b * 2 * a.Z + h], lv = a.lat[b * 2 * a.Z + a.Z + h];
        float dkl_mu, dkl_lv, dsig;
        if (a.kl_form == CRVAE_KL_LOGSIGMA) {
            dkl_mu = mu * invB; dkl_lv = -(1.f - expf(2.f * lv)) * invB; dsig = 0.5f * expf(lv);
        } else if (a.kl_form == CRVAE_KL_SWAPPED) {
            dkl_mu = -0.5f * (1.f - expf(mu)) * invB; dkl_lv = lv * invB; dsig = 0.5f * expf(0.5f * lv);
        } else {
            dkl_mu = mu * invB; dkl_lv = -0.5f * (1.f - expf(lv)) * invB; dsig = 0.5f * expf(0.5f * lv);
        }
        a.dlat[b * 2 * a.Z + h] = dzv[j] + a.beta * dkl_mu;
        a.dlat[b * 2 * a.Z + a.Z + h] = dzv[j] * a.eps[el] * dsig + a.beta * dkl_lv;
    }
}

}  // namespace crvae

using namespace crvae;

extern "C" size_t crvae_dz_allreduce_bytes(int B, int Z, int world) {
    const long long n = (long long)B * Z;
    const long long nchunks = (n + DZ_CHUNK - 1) / DZ_CHUNK;
    return (size_t)(2 * n + (long long)world * nchunks + nchunks) * 4;
}

extern "C" int crvae_dz_allreduce_latent_bwd(const float* dh0, int P, void* const* peer_bufs, int rank, int world, const float* lat,
                                             const float* eps, float beta, int kl_form, float* dlat, float* dz_out, int B, int Z,
                                             void* stream) {
    CRVAE_REQUIRE(peer_bufs && world >= 1 && world <= DZ_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world");
    CRVAE_REQUIRE(B > 0 && Z > 0 && ((long long)B * Z) % 4 == 0 && P >= 0 && (P == 0 || dh0), "bad size");
    CRVAE_REQUIRE(dlat == nullptr || (lat && eps), "lat/eps required with dlat");
    DzArgs a;
    a.dh0 = dh0; a.P = P; a.rank = rank; a.world = world; a.lat = lat; a.eps = eps; a.beta = beta; a.kl_form = kl_form;
    a.dlat = dlat; a.dz_out = dz_out; a.B = B; a.Z = Z; a.n = B * Z; a.nchunks = (a.n + DZ_CHUNK - 1) / DZ_CHUNK;
    for (int r = 0; r < DZ_MAX_WORLD; ++r) a.peers[r] = r < world ? static_cast<float*>(peer_bufs[r]) : nullptr;
    for (int r = 0; r < world; ++r) CRVAE_REQUIRE(a.peers[r] && aligned16(a.peers[r]), "peer buffer missing / misaligned");
    CRVAE_REQUIRE(P == 0 || aligned16(dh0), "16-byte alignment");
    dz_allreduce_latent_bwd_kernel<<<a.nchunks, DZ_THREADS, 0, (cudaStream_t)stream>>>(a);
    return check_launch("dz_allreduce_latent_bwd_kernel");
}

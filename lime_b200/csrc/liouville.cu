// Liouville-space linear ODE  dv/dt = R v  with R an arbitrary CSR matrix: the form in
// which lime propagates the Redfield equation (R from redfield_tensor, lime/oqs.py:528-579;
// loop lime/oqs.py:453-462 with rhs = R.dot(v), lime/oqs.py:471-472) and builds e^{Rt}
// column by column (expm 'EOM', lime/phys.py:1384-1401).
//
// One "slot" of tps threads per vector, several slots per CTA for small D; the vector, the
// RK4 accumulator (registers) and the two stage vectors (shared memory) stay on chip for
// all nsteps steps.
#include "../../include/lime_b200.h"
#include "common.cuh"
#include <cstdlib>

namespace {

struct LvArgs {
    const int* indptr; const int* indices; const cplx* data;
    int D, B, E, nsteps, traj_every, slots, tps;
    cplx* v; const cplx* e; cplx* obs; cplx* traj;
    double dt;
};

template <int EPT>
__global__ void __launch_bounds__(1024, 1)
liouville_rk4_kernel(LvArgs a) {
    extern __shared__ double2 smem[];
    const int D = a.D, tps = a.tps;
    const int slot = threadIdx.x / tps;
    const int t = threadIdx.x - slot * tps;
    const int b = blockIdx.x * a.slots + slot;
    const bool active = b < a.B;
    cplx* y0 = smem + (size_t)slot * 2 * D;
    cplx* y1 = y0 + D;
    cplx* red = smem + (size_t)a.slots * 2 * D;

    cplx v[EPT], acc[EPT];
    bool ok[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        int r = t + e * tps;
        ok[e] = r < D;
        v[e] = (ok[e] && active) ? a.v[(size_t)b * D + r] : cmake(0, 0);
        acc[e] = cmake(0, 0);
        if (ok[e]) y0[r] = v[e];
    }
    __syncthreads();
    const double dt = a.dt, hdt = 0.5 * a.dt;
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const cplx* yin = (stage & 1) ? y1 : y0;
            cplx* yout = (stage & 1) ? y0 : y1;
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                if (!ok[e]) continue;
                const int r = t + e * tps;
                cplx k = cmake(0, 0);
                const int p1 = __ldg(a.indptr + r + 1);
                for (int p = __ldg(a.indptr + r); p < p1; ++p)
                    cfma(k, __ldg(a.data + p), yin[__ldg(a.indices + p)]);
                cplx yn;
                if (stage == 0) {
                    acc[e] = k;
                    yn = cmake(fma(hdt, k.x, v[e].x), fma(hdt, k.y, v[e].y));
                } else if (stage == 1) {
                    rfma(acc[e], 2.0, k);
                    yn = cmake(fma(hdt, k.x, v[e].x), fma(hdt, k.y, v[e].y));
                } else if (stage == 2) {
                    rfma(acc[e], 2.0, k);
                    yn = cmake(fma(dt, k.x, v[e].x), fma(dt, k.y, v[e].y));
                } else {
                    cplx tot = cadd(acc[e], k);
                    v[e].x += tot.x / 6.0 * dt;
                    v[e].y += tot.y / 6.0 * dt;
                    yn = v[e];
                }
                yout[r] = yn;
            }
            __syncthreads();
        }
        if (a.obs) {
            for (int eo = 0; eo < a.E; ++eo) {
                cplx s = cmake(0, 0);
#pragma unroll
                for (int e = 0; e < EPT; ++e)
                    if (ok[e]) cfma(s, __ldg(a.e + (size_t)eo * D + t + e * tps), v[e]);
                if (tps >= 32) {
                    for (int off = 16; off > 0; off >>= 1) {
                        s.x += __shfl_down_sync(0xffffffffu, s.x, off);
                        s.y += __shfl_down_sync(0xffffffffu, s.y, off);
                    }
                    if ((t & 31) == 0) red[slot * 32 + (t >> 5)] = s;
                    __syncthreads();
                    if (t == 0 && active) {
                        cplx tot = cmake(0, 0);
                        for (int w = 0; w < tps / 32; ++w) tot = cadd(tot, red[slot * 32 + w]);
                        a.obs[((size_t)step * a.B + b) * a.E + eo] = tot;
                    }
                    __syncthreads();
                } else {
                    for (int off = tps >> 1; off > 0; off >>= 1) {
                        s.x += __shfl_xor_sync(0xffffffffu, s.x, off);
                        s.y += __shfl_xor_sync(0xffffffffu, s.y, off);
                    }
                    if (t == 0 && active) a.obs[((size_t)step * a.B + b) * a.E + eo] = s;
                }
            }
        }
        if (a.traj && ((step + 1) % a.traj_every) == 0 && active) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * D;
#pragma unroll
            for (int e = 0; e < EPT; ++e)
                if (ok[e]) dst[t + e * tps] = v[e];
        }
    }
    if (active) {
#pragma unroll
        for (int e = 0; e < EPT; ++e)
            if (ok[e]) a.v[(size_t)b * D + t + e * tps] = v[e];
    }
}

// Small Liouville dimension (instantiated for D = 4, the two-level system): ONE THREAD PER VECTOR.  The vector, the RK4
// accumulator and the stage vector live in registers; the densified generator and the observable rows sit in
// shared memory and are read with warp-uniform (broadcast) loads; no barrier inside the time loop.
template <int D>
__global__ void __launch_bounds__(128)
liouville_small_kernel(LvArgs a) {
    extern __shared__ double2 smem[];
    cplx* Rs = smem;                    // [D][D] dense, row-major
    cplx* es = smem + D * D;            // [E][D]
    for (int l = threadIdx.x; l < D * D; l += blockDim.x) Rs[l] = cmake(0, 0);
    for (int l = threadIdx.x; l < a.E * D; l += blockDim.x) es[l] = a.e[l];
    __syncthreads();
    for (int r = threadIdx.x; r < D; r += blockDim.x)
        for (int p = a.indptr[r]; p < a.indptr[r + 1]; ++p) {
            cplx& d = Rs[r * D + a.indices[p]];
            d = cadd(d, a.data[p]);
        }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    cplx v[D], acc[D], y[D];
    const cplx* gv = a.v + (size_t)b * D;
#pragma unroll
    for (int r = 0; r < D; ++r) { v[r] = gv[r]; y[r] = v[r]; }
    const double dt = a.dt, hdt = 0.5 * a.dt, w6 = a.dt / 6.0;
    for (int step = 0; step < a.nsteps; ++step) {
#pragma unroll
        for (int stage = 0; stage < 4; ++stage) {
            cplx k[D];
#pragma unroll
            for (int r = 0; r < D; ++r) {
                cplx s = cmake(0, 0);
#pragma unroll
                for (int c = 0; c < D; ++c) cfma(s, Rs[r * D + c], y[c]);
                k[r] = s;
            }
#pragma unroll
            for (int r = 0; r < D; ++r) {
                if (stage == 0) {
                    acc[r] = k[r];
                    y[r] = cmake(fma(hdt, k[r].x, v[r].x), fma(hdt, k[r].y, v[r].y));
                } else if (stage == 1) {
                    rfma(acc[r], 2.0, k[r]);
                    y[r] = cmake(fma(hdt, k[r].x, v[r].x), fma(hdt, k[r].y, v[r].y));
                } else if (stage == 2) {
                    rfma(acc[r], 2.0, k[r]);
                    y[r] = cmake(fma(dt, k[r].x, v[r].x), fma(dt, k[r].y, v[r].y));
                } else {
                    v[r].x = fma(w6, acc[r].x + k[r].x, v[r].x);
                    v[r].y = fma(w6, acc[r].y + k[r].y, v[r].y);
                    y[r] = v[r];
                }
            }
        }
        if (a.obs)
            for (int eo = 0; eo < a.E; ++eo) {
                cplx s = cmake(0, 0);
#pragma unroll
                for (int r = 0; r < D; ++r) cfma(s, es[eo * D + r], v[r]);
                a.obs[((size_t)step * a.B + b) * a.E + eo] = s;
            }
        if (a.traj && ((step + 1) % a.traj_every) == 0) {
            cplx* dst = a.traj + ((size_t)(step / a.traj_every) * a.B + b) * D;
#pragma unroll
            for (int r = 0; r < D; ++r) dst[r] = v[r];
        }
    }
    cplx* out = a.v + (size_t)b * D;
#pragma unroll
    for (int r = 0; r < D; ++r) out[r] = v[r];
}

}  // namespace

extern "C" int limeb200_liouville_rk4_csr(const int* d_indptr, const int* d_indices, const double* d_data,
                                          int D, double* d_v, int B, const double* d_e, int E,
                                          double* d_obs, double* d_traj, int traj_every,
                                          double dt, int nsteps, void* stream) {
    LB_REQUIRE(d_indptr && d_indices && d_data && d_v, "null argument");
    LB_REQUIRE(D >= 1 && D <= 4096, "Liouville dimension D=%d out of range (1..4096); use the operator-form plan", D);
    LB_REQUIRE(B >= 1 && nsteps >= 0 && E >= 0, "bad arguments");
    LB_REQUIRE(E == 0 || (d_e && d_obs), "observables requested without buffers");
    LB_REQUIRE(!d_traj || traj_every >= 1, "traj_every must be >= 1");
    if (nsteps == 0) return LB_OK;
    int dev = 0;
    LB_CUDA(cudaGetDevice(&dev));
    int sms = 148;
    LB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LvArgs a;
    a.indptr = d_indptr; a.indices = d_indices; a.data = (const cplx*)d_data;
    a.D = D; a.B = B; a.E = E; a.nsteps = nsteps; a.traj_every = d_traj ? traj_every : 1;
    a.v = (cplx*)d_v; a.e = (const cplx*)d_e; a.obs = E > 0 ? (cplx*)d_obs : nullptr; a.traj = (cplx*)d_traj;
    a.dt = dt;
    if (D == 4 && E <= 16 && B >= 4096 && !getenv("LIMEB200_LIOUVILLE_GENERIC")) {
        // thread-per-vector kernel for batches of two-level systems (D = 9, 16 would spill: 8 D doubles of state)
        const size_t smem = (size_t)(D * D + E * D) * 16;
        void (*k)(LvArgs) = liouville_small_kernel<4>;
        k<<<ceil_div(B, 128), 128, smem, (cudaStream_t)stream>>>(a);
        LB_CUDA(cudaGetLastError());
        return LB_OK;
    }
    int ept = 1, tps;
    if (D <= 4) tps = 4;
    else if (D <= 8) tps = 8;
    else if (D <= 16) tps = 16;
    else {
        if (D > 2048) ept = 4; else if (D > 1024) ept = 2;
        tps = ceil_div(ceil_div(D, ept), 32) * 32;
    }
    int slots = std::max(1, 256 / tps);
    slots = std::min(slots, B);
    while (slots > 1 && ceil_div(B, slots) < 2 * sms) slots >>= 1;
    a.slots = slots; a.tps = tps;
    size_t smem = (size_t)slots * (2 * D + 32) * 16;
    void (*kern)(LvArgs) = ept == 1 ? liouville_rk4_kernel<1> : (ept == 2 ? liouville_rk4_kernel<2> : liouville_rk4_kernel<4>);
    LB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ceil_div(B, slots), slots * tps, smem, (cudaStream_t)stream>>>(a);
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

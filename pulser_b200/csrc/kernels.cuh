// sm_100a kernels of the matrix-free propagator.
//
// Hot op (one Clenshaw stage of the Chebyshev expansion of exp(-iG)):
//     out[s] = c_psi*psi[s] + c_b2*b2[s] + c_g * (Gt v)[s]
//     (Gt v)[s] = (w*Dint[s] - sum_k th_k [digit_k(s)==from] - gamma) v[s]
//               + sum_k ( digit_k(s)==to ? g_k : conj(g_k) ) v[s with digit_k swapped]
// which restates, matrix-free, the CSR products QuTiP performs for the QobjEvo
// built at pulser-simulation/pulser_simulation/hamiltonian.py:246-439
// (SURVEY.md Appendix A.3).  HBM/L2-bound: algorithmic traffic is
// 16 (v) + 8 (Dint) + 16 (out) = 40 B per amplitude per apply.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb200 {

struct c2 { double x, y; };

__host__ __device__ inline c2 cmul(c2 a, c2 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__host__ __device__ inline c2 cadd(c2 a, c2 b) { return {a.x + b.x, a.y + b.y}; }

// ---- tile geometry of one pass -------------------------------------------
// A tile gathers the amplitudes whose index differs only in the bits
// [0, lo_bits) and [hi_shift, hi_shift + hi_bits); it is closed under flips of
// those bits, so their partners are served from shared memory.  Bits in
// `extra_mask` are flipped through coalesced global loads.
struct PassGeom {
    int n_bits;      // N (d = 2)
    int lo_bits;     // contiguous low bits in the tile (row = 2^lo_bits amps)
    int hi_shift;    // first bit of the high group
    int hi_bits;     // bits in the high group (rows = 2^hi_bits)
    uint32_t tile_flip_mask;   // tile-local bit positions whose flips belong to this pass
    unsigned long long extra_mask;  // global bit positions flipped via global loads
    int first_pass;  // 1: psi, b2 and the diagonal are added in this pass
};

struct StageCoef {  // complex scalars of the Clenshaw stage
    c2 c_psi, c_b2, c_g;
};

struct UniformDrive {  // same coefficients on every qubit and trajectory
    c2 g;          // scaled drive  g/rho
    double theta;  // scaled detuning moment
    double w;      // scaled weight of Dint
    double gamma;  // scaled centre
    int to_bit;    // digit value of |to> (1 for ground-rydberg / digital)
    int from_count_is_popc;  // 1: [digit==from] counted by popc(idx) (from digit = 1)
};

// per-(exponential, trajectory) table for non-uniform drives, d = 2:
//   tab[0 .. 2N)      g (re, im) per BIT position p
//   tab[2N .. 3N)     theta per bit position p
//   tab[3N], tab[3N+1] w, gamma
__host__ __device__ inline int d2_table_stride(int n) { return 3 * n + 2; }

// ---- PTX helpers: mbarrier + TMA 1-D bulk copy ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// streaming (read-once) 16-byte load / store
__device__ __forceinline__ c2 ld_stream(const c2* p) {
    double2 r = __ldcs(reinterpret_cast<const double2*>(p));
    return {r.x, r.y};
}
__device__ __forceinline__ void st_c2(c2* p, c2 v) {
    *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
}

// Fused Lanczos step (StageArgs::lz set).  The gather source `v` holds the RAW vector r_j = G v_j - beta_{j-1} v_{j-1}
// of the previous stage; alpha_j = Re<v_j, r_j> and |r_j|^2 were reduced by that stage into lz.acc_prev.  This stage
//   v_{j+1} = (r_j - alpha_j v_j) / beta_j                           (second output, own element)
//   r_{j+1} = G v_{j+1} - beta_j v_j
//           = [G r_j - alpha_j r_j - alpha_j beta_{j-1} v_{j-1}] / beta_j - beta_j v_j     (G v_j = r_j + beta_{j-1} v_{j-1})
// so the separate vector-update kernel (48 B/amplitude, one launch per iteration) disappears: 88 B/amplitude per
// Lanczos iteration instead of 104, one launch.  Reference call replaced: qutip.sesolve, simulation.py:729-735.
struct LanczosFuse {
    const c2* vj;            // v_j      [B][D] own element
    const c2* vjm1;          // v_{j-1}  [B][D] own element (nullptr for j = 0)
    c2* vout;                // v_{j+1}  [B][D]
    const double* acc_prev;  // [B][2]: alpha_j, |r_j|^2
    const double* beta_prev; // [B]: beta_{j-1} (nullptr for j = 0)
    double* alpha_out;       // [B]: alpha_j recorded for the host
    double* beta_out;        // [B]: beta_j
    double* acc_clear;       // [B][2]: accumulator of the stage after this one, cleared here
};

struct LanczosCoef { double alpha, beta, inv, beta_prev; };

__device__ __forceinline__ LanczosCoef lanczos_coef(const LanczosFuse& lz, long long traj) {
    LanczosCoef c;
    c.alpha = lz.acc_prev[2 * traj];
    const double ww = lz.acc_prev[2 * traj + 1];
    const double b2 = ww - c.alpha * c.alpha;
    c.beta = (b2 > 1e-28 * fmax(ww, 1e-300)) ? sqrt(b2) : 0.0;   // 0: breakdown (invariant subspace reached)
    c.inv = c.beta > 0.0 ? 1.0 / c.beta : 0.0;
    c.beta_prev = lz.beta_prev ? lz.beta_prev[traj] : 0.0;
    return c;
}

// ---- d = 2 tiled stage kernel ---------------------------------------------
struct StageArgs {
    const c2* v;      // gather source      [B][D]
    const c2* psi;    // own element        [B][D]
    const c2* b2;     // own element        [B][D] (may alias out)
    c2* out;          // [B][D]
    const double* dint;        // [Bd][D]
    long long dint_stride;     // 0 when shared by all trajectories
    long long D;               // 2^N
    PassGeom geo;
    StageCoef coef;
    UniformDrive u;            // used when UNIFORM
    const double* table;       // [B][stride] for this exponential (non-uniform)
    int to_bit;
    int from_is_one;
    const double* beta_dev;  // Lanczos: c_b2 = -beta_dev[traj] read on the device (nullptr: use coef.c_b2)
    double* dot_acc;         // Lanczos: if set, acc[traj][0] += Re<lhs, out>, acc[traj][1] += <out, out> (fused reductions;
                             // lhs = v, or v_{j+1} in a fused Lanczos step)
    LanczosFuse lz;          // fused Lanczos step when lz.vj != nullptr (register-blocked kernels)
    // partner-sum forwarding (stage_d2_fwd_kernel, uniform drives): w_in[s] = sum of v over the flips this stage
    // does NOT perform (its producer's tile was closed under them), w_out[s] = the same sum of `out` over THIS
    // stage's tile flips for the consumer.  Plane 0 ([D]) holds P = sum v[s^k]; plane 1 (at + D * n_traj) holds the
    // signed sum Q a complex drive also needs.  Both may be null (first stage of a chain / nobody follows).
    const c2* w_in;
    c2* w_out;
    long long w_plane;       // distance between the P and the Q plane
};

// up to two independent Clenshaw chains per launch (the h and the h/2 branches of a Richardson step):
// blockIdx.y = chain * n_traj + trajectory
struct StageArgs2 {
    StageArgs a[2];
    int n_traj;
};

template <bool UNIFORM, bool REAL_G>
__global__ void __launch_bounds__(256) stage_d2_kernel(StageArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c2* tile = reinterpret_cast<c2*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;

    const PassGeom g = a.geo;
    const int tbits = g.lo_bits + g.hi_bits;
    const int tsize = 1 << tbits;
    const long long traj = blockIdx.y;
    const long long tile_id = blockIdx.x;
    // bits of tile_id fill positions [lo, hi_shift) and [hi_shift + hi_bits, N)
    const int mid_bits = g.hi_shift - g.lo_bits;
    const long long mid = tile_id & ((1LL << mid_bits) - 1);
    const long long top = tile_id >> mid_bits;
    const long long base = (mid << g.lo_bits) | (top << (g.hi_shift + g.hi_bits));
    const long long voff = traj * a.D;
    const c2* vsrc = a.v + voff;

    // per-bit tables for non-uniform drives live after the tile
    double* tab = reinterpret_cast<double*>(tile + tsize);
    if (!UNIFORM) {
        const int stride = d2_table_stride(g.n_bits);
        const double* src = a.table + traj * stride;
        for (int i = threadIdx.x; i < stride; i += blockDim.x) tab[i] = src[i];
    }

    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&mbar, (uint32_t)tsize * 16u);
    {
        const int rows = 1 << g.hi_bits;
        const uint32_t row_bytes = (uint32_t)(16u << g.lo_bits);
        for (int r = threadIdx.x; r < rows; r += blockDim.x)
            tma_load_1d(tile + ((size_t)r << g.lo_bits), vsrc + base + ((long long)r << g.hi_shift), row_bytes, &mbar);
    }
    mbar_wait(&mbar, 0);

    const long long lomask = (1LL << g.lo_bits) - 1;
    double w, gamma, theta_u;
    c2 gu;
    if (UNIFORM) {
        w = a.u.w; gamma = a.u.gamma; theta_u = a.u.theta; gu = a.u.g;
    } else {
        w = tab[3 * g.n_bits]; gamma = tab[3 * g.n_bits + 1]; theta_u = 0.0; gu = {0.0, 0.0};
    }
    const int to_bit = a.to_bit;

    for (int t = threadIdx.x; t < tsize; t += blockDim.x) {
        const long long idx = base | (t & lomask) | ((long long)(t >> g.lo_bits) << g.hi_shift);
        const c2 vo = tile[t];
        double pr = 0.0, pi = 0.0, qr = 0.0, qi = 0.0;  // uniform: P, Q sums; non-uniform: pr,pi = drive result
        // flips served from shared memory
#pragma unroll 1
        for (uint32_t m = g.tile_flip_mask; m; m &= m - 1) {
            const int j = __ffs(m) - 1;
            const c2 pv = tile[t ^ (1 << j)];
            const int bit = (t >> j) & 1;
            if (UNIFORM) {
                pr += pv.x; pi += pv.y;
                if (!REAL_G) {
                    const double s = (bit == to_bit) ? 1.0 : -1.0;
                    qr = fma(s, pv.x, qr); qi = fma(s, pv.y, qi);
                }
            } else {
                const int p = (j < g.lo_bits) ? j : (j - g.lo_bits + g.hi_shift);
                const double gx = tab[2 * p];
                const double gy = (bit == to_bit) ? tab[2 * p + 1] : -tab[2 * p + 1];
                pr = fma(gx, pv.x, pr); pr = fma(-gy, pv.y, pr);
                pi = fma(gx, pv.y, pi); pi = fma(gy, pv.x, pi);
            }
        }
        // flips served by coalesced global loads
#pragma unroll 1
        for (unsigned long long m = g.extra_mask; m; m &= m - 1) {
            const int p = __ffsll((long long)m) - 1;
            const double2 raw = __ldg(reinterpret_cast<const double2*>(vsrc + (idx ^ (1LL << p))));
            const c2 pv = {raw.x, raw.y};
            const int bit = (int)((idx >> p) & 1);
            if (UNIFORM) {
                pr += pv.x; pi += pv.y;
                if (!REAL_G) {
                    const double s = (bit == to_bit) ? 1.0 : -1.0;
                    qr = fma(s, pv.x, qr); qi = fma(s, pv.y, qi);
                }
            } else {
                const double gx = tab[2 * p];
                const double gy = (bit == to_bit) ? tab[2 * p + 1] : -tab[2 * p + 1];
                pr = fma(gx, pv.x, pr); pr = fma(-gy, pv.y, pr);
                pi = fma(gx, pv.y, pi); pi = fma(gy, pv.x, pi);
            }
        }
        c2 drive;
        if (UNIFORM) {
            // g*S_to + conj(g)*S_from = x*P + i*y*Q
            drive.x = gu.x * pr; drive.y = gu.x * pi;
            if (!REAL_G) { drive.x = fma(-gu.y, qi, drive.x); drive.y = fma(gu.y, qr, drive.y); }
        } else {
            drive = {pr, pi};
        }
        c2 res;
        if (g.first_pass) {
            double diag = -gamma;
            if (a.dint) diag = fma(w, __ldcs(a.dint + traj * a.dint_stride + idx), diag);
            if (UNIFORM) {
                const int ones = __popcll((unsigned long long)idx);
                const int cnt = a.from_is_one ? ones : (g.n_bits - ones);
                diag = fma(-theta_u, (double)cnt, diag);
            } else {
                double acc = 0.0;
                for (int p = 0; p < g.n_bits; ++p) {
                    const int bit = (int)((idx >> p) & 1);
                    acc += (bit == a.from_is_one) ? tab[2 * g.n_bits + p] : 0.0;
                }
                diag -= acc;
            }
            c2 gv = {fma(diag, vo.x, drive.x), fma(diag, vo.y, drive.y)};
            res = cmul(a.coef.c_g, gv);
            if (a.psi) res = cadd(res, cmul(a.coef.c_psi, ld_stream(a.psi + voff + idx)));
            if (a.b2) res = cadd(res, cmul(a.beta_dev ? c2{-a.beta_dev[traj], 0.0} : a.coef.c_b2, ld_stream(a.b2 + voff + idx)));
        } else {
            res = cadd(ld_stream(a.out + voff + idx), cmul(a.coef.c_g, drive));
        }
        st_c2(a.out + voff + idx, res);
    }
}

// ---- d = 2 tiled stage kernel, register-blocked (the production kernel) ------
// Each thread owns R = 2^RB amplitudes of the tile (tile index t = tid + r*NT,
// i.e. the top RB tile bits live in registers): flips of those bits are
// register-to-register, flips of the other tile bits cost one LDS.128 per owned
// amplitude, all independent (fully unrolled) so that the shared-memory pipe
// stays full.  Shared-memory operand traffic per amplitude and pass is
// (flipped tile bits - RB + 1) x 16 B.
// Compute phase of one tile, shared by the one-shot and the persistent kernels: gathers from the tile in
// shared memory (register-blocked), optional global-load partners, fused epilogue and store.
// own-element global load: streaming (read once per stage)
__device__ __forceinline__ c2 ld_own(const c2* p) {
    double2 r = __ldcs(reinterpret_cast<const double2*>(p));
    return {r.x, r.y};
}

// In-tile partner sums of the R = 2^RB amplitudes a thread owns (tile index t = tid + r*NT): flips of the
// register-block bits (tile bits TBITS-RB .. TBITS-1) are register moves, flips of the tile bits
// [jstart, TBITS-RB) are independent LDS.128 from `tile`.
template <bool UNIFORM, bool REAL_G, int TBITS, int RB>
__device__ __forceinline__ void rb_tile_gather(const PassGeom& g, const c2* tile, const double* __restrict__ tab, int tid,
                                               int to_bit, int jstart, bool skip_smem, const c2 (&v)[1 << RB],
                                               double (&pr)[1 << RB], double (&pi)[1 << RB], double (&qr)[1 << RB],
                                               double (&qi)[1 << RB]) {
    constexpr int R = 1 << RB;
    constexpr int NT = 1 << (TBITS - RB);
    // --- flips inside the register block (tile bits TBITS-RB .. TBITS-1) ---
#pragma unroll
    for (int q = 0; q < RB; ++q) {
        const int j = TBITS - RB + q;
        double gx = 0.0, gyt = 0.0;
        if (!UNIFORM) {
            const int p = (j < g.lo_bits) ? j : (j - g.lo_bits + g.hi_shift);
            gx = tab[2 * p]; gyt = tab[2 * p + 1];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const c2 pv = v[r ^ (1 << q)];
            const int bit = (r >> q) & 1;
            if (UNIFORM) {
                pr[r] += pv.x; pi[r] += pv.y;
                if (!REAL_G) {
                    if (bit == to_bit) { qr[r] += pv.x; qi[r] += pv.y; } else { qr[r] -= pv.x; qi[r] -= pv.y; }
                }
            } else {
                const double gy = (bit == to_bit) ? gyt : -gyt;
                pr[r] = fma(gx, pv.x, pr[r]); pr[r] = fma(-gy, pv.y, pr[r]);
                pi[r] = fma(gx, pv.y, pi[r]); pi[r] = fma(gy, pv.x, pi[r]);
            }
        }
    }
    // --- flips served from shared memory ---
#pragma unroll
    for (int j = 0; j < TBITS - RB; ++j) {
        if (j >= jstart && !skip_smem) {
            const int bit = (tid >> j) & 1;
            const int ptid = tid ^ (1 << j);
            double gx = 0.0, gy = 0.0;
            if (!UNIFORM) {
                const int p = (j < g.lo_bits) ? j : (j - g.lo_bits + g.hi_shift);
                gx = tab[2 * p];
                gy = (bit == to_bit) ? tab[2 * p + 1] : -tab[2 * p + 1];
            }
            const double sg = (bit == to_bit) ? 1.0 : -1.0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const c2 pv = tile[ptid + r * NT];
                if (UNIFORM) {
                    pr[r] += pv.x; pi[r] += pv.y;
                    if (!REAL_G) { qr[r] = fma(sg, pv.x, qr[r]); qi[r] = fma(sg, pv.y, qi[r]); }
                } else {
                    pr[r] = fma(gx, pv.x, pr[r]); pr[r] = fma(-gy, pv.y, pr[r]);
                    pi[r] = fma(gx, pv.y, pi[r]); pi[r] = fma(gy, pv.x, pi[r]);
                }
            }
        }
    }
}

// Compute phase of one tile: gathers from the tile in shared memory (register-blocked), coalesced global loads
// for the partners outside the tile, fused epilogue (diagonal, Clenshaw / Lanczos combination, reductions) and store.
template <bool UNIFORM, bool REAL_G, int TBITS, int RB>
__device__ __forceinline__ void rb_tile_compute(const StageArgs& a, const PassGeom& g, const c2* tile,
                                                const double* __restrict__ tab, long long base, long long traj,
                                                int tid, uint64_t* tile_bar) {
    constexpr int R = 1 << RB;
    constexpr int NT = 1 << (TBITS - RB);
    const long long voff = traj * a.D;
    const c2* vsrc = a.v + voff;
    const long long lomask = (1LL << g.lo_bits) - 1;
    const int to_bit = a.to_bit;
    // first tile bit whose flip belongs to this pass (pass A: 0, later passes: lo_bits)
    const int jstart = __ffs(g.tile_flip_mask) - 1;

    c2 v[R];
    double pr[R], pi[R], qr[R], qi[R];
    // global index of each owned amplitude
    long long idx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int t = tid + r * NT;
        idx[r] = base | (t & lomask) | ((long long)(t >> g.lo_bits) << g.hi_shift);
        pr[r] = 0.0; pi[r] = 0.0; qr[r] = 0.0; qi[r] = 0.0;
    }
    // --- flips of the bits outside the tile: coalesced partner loads.  They do not depend on the tile, so they go
    //     out while the bulk copy of the tile is still in flight (the wait on its mbarrier comes after them) ---
    for (unsigned long long m = g.extra_mask; m; m &= m - 1) {
        const int p = __ffsll((long long)m) - 1;
        double gx = 0.0, gyt = 0.0;
        if (!UNIFORM) { gx = tab[2 * p]; gyt = tab[2 * p + 1]; }
        const int bit = (int)((base >> p) & 1);  // extra bits are never tile bits
        const double sg = (bit == to_bit) ? 1.0 : -1.0;
        const double gy = (bit == to_bit) ? gyt : -gyt;
        double2 raw[R];
#pragma unroll
        for (int r = 0; r < R; ++r) raw[r] = __ldg(reinterpret_cast<const double2*>(vsrc + (idx[r] ^ (1LL << p))));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (UNIFORM) {
                pr[r] += raw[r].x; pi[r] += raw[r].y;
                if (!REAL_G) { qr[r] = fma(sg, raw[r].x, qr[r]); qi[r] = fma(sg, raw[r].y, qi[r]); }
            } else {
                pr[r] = fma(gx, raw[r].x, pr[r]); pr[r] = fma(-gy, raw[r].y, pr[r]);
                pi[r] = fma(gx, raw[r].y, pi[r]); pi[r] = fma(gy, raw[r].x, pi[r]);
            }
        }
    }
    if (tile_bar) mbar_wait(tile_bar, 0);
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = tile[tid + r * NT];
    rb_tile_gather<UNIFORM, REAL_G, TBITS, RB>(g, tile, tab, tid, to_bit, jstart, false, v, pr, pi, qr, qi);
    // --- epilogue ---
    double w = 0.0, gamma = 0.0, th_common = 0.0;
    double th_r[R];
    if (g.first_pass) {
        if (UNIFORM) { w = a.u.w; gamma = a.u.gamma; }
        else {
            w = tab[3 * g.n_bits]; gamma = tab[3 * g.n_bits + 1];
            // theta sum split into (bits of base) + (bits of tid) + (register bits)
            const long long fixed = base | (tid & lomask) | ((long long)(tid >> g.lo_bits) << g.hi_shift);
            for (int p = 0; p < g.n_bits; ++p) {
                const int bit = (int)((fixed >> p) & 1);
                th_common += (bit == a.from_is_one) ? tab[2 * g.n_bits + p] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < RB; ++q) {
                    const int j = TBITS - RB + q;
                    const int p = (j < g.lo_bits) ? j : (j - g.lo_bits + g.hi_shift);
                    // `fixed` has these bits at 0: replace the bit-0 contribution by the bit-(r>>q) one
                    const double th = tab[2 * g.n_bits + p];
                    const int bit = (r >> q) & 1;
                    acc += ((bit == a.from_is_one) ? th : 0.0) - ((0 == a.from_is_one) ? th : 0.0);
                }
                th_r[r] = acc;
            }
        }
    }
    // drive term of every owned amplitude (frees the P/Q accumulators)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (UNIFORM) {
            const double dx = a.u.g.x * pr[r], dy = a.u.g.x * pi[r];
            if (!REAL_G) { pr[r] = fma(-a.u.g.y, qi[r], dx); pi[r] = fma(a.u.g.y, qr[r], dy); }
            else { pr[r] = dx; pi[r] = dy; }
        }
    }
    // own-element global operands are staged in registers half a block at a time so that the loads
    // of one half are all in flight together (the stores to `out` may alias them for the compiler)
    constexpr int H = (R >= 4) ? R / 2 : R;
    const c2 cb2 = a.beta_dev ? c2{-a.beta_dev[traj], 0.0} : a.coef.c_b2;
    const bool fuse = a.lz.vj != nullptr;
    LanczosCoef lc{0.0, 0.0, 0.0, 0.0};
    if (fuse) lc = lanczos_coef(a.lz, traj);
    double dot0 = 0.0, dot1 = 0.0;
    if (g.first_pass) {
        const double* dsrc = a.dint ? a.dint + traj * a.dint_stride : nullptr;
#pragma unroll
        for (int h0 = 0; h0 < R; h0 += H) {
            double dv[H];
            c2 pv[H], bv[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                dv[r] = dsrc ? __ldcs(dsrc + idx[h0 + r]) : 0.0;
                if (fuse) {
                    pv[r] = ld_own(a.lz.vj + voff + idx[h0 + r]);
                    bv[r] = (a.lz.vjm1 && lc.beta_prev != 0.0) ? ld_own(a.lz.vjm1 + voff + idx[h0 + r]) : c2{0.0, 0.0};
                } else {
                    pv[r] = a.psi ? ld_own(a.psi + voff + idx[h0 + r]) : c2{0.0, 0.0};
                    bv[r] = a.b2 ? ld_own(a.b2 + voff + idx[h0 + r]) : c2{0.0, 0.0};
                }
            }
#pragma unroll
            for (int r = 0; r < H; ++r) {
                const int rr = h0 + r;
                double diag = fma(w, dv[r], -gamma);
                if (UNIFORM) {
                    const int ones = __popcll((unsigned long long)idx[rr]);
                    const int cnt = a.from_is_one ? ones : (g.n_bits - ones);
                    diag = fma(-a.u.theta, (double)cnt, diag);
                } else {
                    diag -= th_common + th_r[rr];
                }
                const c2 gv = {fma(diag, v[rr].x, pr[rr]), fma(diag, v[rr].y, pi[rr])};
                c2 res, lhs;
                if (fuse) {
                    // v_{j+1} and r_{j+1} from the raw vector (see LanczosFuse)
                    const c2 vn = {(v[rr].x - lc.alpha * pv[r].x) * lc.inv, (v[rr].y - lc.alpha * pv[r].y) * lc.inv};
                    const double ai = lc.alpha * lc.inv, abi = ai * lc.beta_prev;
                    res.x = fma(lc.inv, gv.x, -fma(ai, v[rr].x, fma(abi, bv[r].x, lc.beta * pv[r].x)));
                    res.y = fma(lc.inv, gv.y, -fma(ai, v[rr].y, fma(abi, bv[r].y, lc.beta * pv[r].y)));
                    st_c2(a.lz.vout + voff + idx[rr], vn);
                    lhs = vn;
                } else {
                    res = cmul(a.coef.c_g, gv);
                    res = cadd(res, cmul(a.coef.c_psi, pv[r]));
                    res = cadd(res, cmul(cb2, bv[r]));
                    lhs = v[rr];
                }
                dot0 = fma(lhs.x, res.x, dot0); dot0 = fma(lhs.y, res.y, dot0);
                dot1 = fma(res.x, res.x, dot1); dot1 = fma(res.y, res.y, dot1);
                st_c2(a.out + voff + idx[rr], res);
            }
        }
    } else {
#pragma unroll
        for (int h0 = 0; h0 < R; h0 += H) {
            c2 ov[H];
#pragma unroll
            for (int r = 0; r < H; ++r) ov[r] = ld_own(a.out + voff + idx[h0 + r]);
#pragma unroll
            for (int r = 0; r < H; ++r) {
                const int rr = h0 + r;
                const c2 res = cadd(ov[r], cmul(a.coef.c_g, c2{pr[rr], pi[rr]}));
                dot0 = fma(v[rr].x, res.x, dot0); dot0 = fma(v[rr].y, res.y, dot0);
                dot1 = fma(res.x, res.x, dot1); dot1 = fma(res.y, res.y, dot1);
                st_c2(a.out + voff + idx[rr], res);
            }
        }
    }
    if (a.dot_acc) {  // block reduction of the fused Lanczos inner products, one atomic pair per CTA
        for (int o = 16; o > 0; o >>= 1) {
            dot0 += __shfl_xor_sync(0xffffffffu, dot0, o);
            dot1 += __shfl_xor_sync(0xffffffffu, dot1, o);
        }
        __shared__ double dred[2][NT / 32 > 0 ? NT / 32 : 1];
        if ((tid & 31) == 0) { dred[0][tid >> 5] = dot0; dred[1][tid >> 5] = dot1; }
        __syncthreads();
        if (tid == 0) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < NT / 32; ++i) { s0 += dred[0][i]; s1 += dred[1][i]; }
            atomicAdd(a.dot_acc + 2 * traj, s0);
            atomicAdd(a.dot_acc + 2 * traj + 1, s1);
        }
    }
    if (fuse && blockIdx.x == 0 && tid == 0) {  // one CTA per trajectory records the recurrence coefficients
        a.lz.alpha_out[traj] = lc.alpha;
        a.lz.beta_out[traj] = lc.beta;
        a.lz.acc_clear[2 * traj] = 0.0;
        a.lz.acc_clear[2 * traj + 1] = 0.0;
    }
}

__device__ __forceinline__ long long tile_base_of(const PassGeom& g, long long tile_id) {
    // bits of tile_id fill positions [lo, hi_shift) and [hi_shift + hi_bits, N)
    const int mid_bits = g.hi_shift - g.lo_bits;
    const long long mid = tile_id & ((1LL << mid_bits) - 1);
    const long long top = tile_id >> mid_bits;
    return (mid << g.lo_bits) | (top << (g.hi_shift + g.hi_bits));
}

// programmatic dependent launch: wait for the producer grid's memory before the first global read
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- d = 2 tiled stage kernel, one tile per CTA --------------------------------
// Each thread owns R = 2^RB amplitudes of the tile (tile index t = tid + r*NT,
// i.e. the top RB tile bits live in registers): flips of those bits are
// register-to-register, flips of the other tile bits cost one LDS.128 per owned
// amplitude, all independent (fully unrolled) so that the shared-memory pipe
// stays full.  Shared-memory operand traffic per amplitude and pass is
// (flipped tile bits - RB + 1) x 16 B.
template <bool UNIFORM, bool REAL_G, int TBITS, int RB>
__global__ void __launch_bounds__(1 << (TBITS - RB), (65536 / ((1 << (TBITS - RB)) * (RB >= 3 ? 128 : 64))))
stage_d2_rb_kernel(const __grid_constant__ StageArgs2 m) {
    constexpr int NT = 1 << (TBITS - RB);
    constexpr int TSIZE = 1 << TBITS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c2* tile = reinterpret_cast<c2*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;

    const int chain = blockIdx.y / m.n_traj;
    const StageArgs& a = m.a[chain];
    const PassGeom g = a.geo;
    const int tid = threadIdx.x;
    const long long traj = blockIdx.y - chain * m.n_traj;
    const long long base = tile_base_of(g, blockIdx.x);
    const c2* vsrc = a.v + traj * a.D;

    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();          // the barrier is initialised before any thread issues a copy on it
    pdl_wait();
    pdl_launch_dependents();
    // the tile copy goes out first; the per-trajectory coefficient table (non-uniform drives) is read behind it
    if (tid == 0) mbar_arrive_expect_tx(&mbar, (uint32_t)TSIZE * 16u);
    const int rows = 1 << g.hi_bits;
    const uint32_t row_bytes = (uint32_t)(16u << g.lo_bits);
    for (int r = tid; r < rows; r += NT)
        tma_load_1d(tile + ((size_t)r << g.lo_bits), vsrc + base + ((long long)r << g.hi_shift), row_bytes, &mbar);
    double* tab = reinterpret_cast<double*>(tile + TSIZE);
    if (!UNIFORM) {
        const int stride = d2_table_stride(g.n_bits);
        const double* src = a.table + traj * stride;
        for (int i = tid; i < stride; i += NT) tab[i] = src[i];
        __syncthreads();
    }
    rb_tile_compute<UNIFORM, REAL_G, TBITS, RB>(a, g, tile, tab, base, traj, tid, &mbar);
}

// ---- d = 2 stage kernel with partner-sum forwarding (uniform drives) ---------------------------------------------
// The single-pass kernel above sits at the L2 throughput cap (~6300 B/clk chip-wide): (N - TBITS) x 16 B of partner
// loads per amplitude dominate its L2 sectors (profiles/r02_l2_hint_experiment.json).  Here consecutive Clenshaw
// stages alternate between two tile geometries with complementary flip sets -- A: the TBITS low bits; B: the
// hb = min(N - TBITS, TBITS - 2) bits above them, gathered as 2^hb rows of 2^(TBITS - hb) amplitudes -- and a stage
// receives the partner sums over the OTHER geometry's flips from the stage that produced its input (w_in, 16 B per
// amplitude) and emits the sums of its own result over ITS flips (w_out): 104 B of L2 traffic per amplitude and stage
// instead of 72 + 16 (N - TBITS).  Bits above TBITS + hb (N > 20) stay coalesced partner loads in both geometries.
// Unlike the round-1 attempt (latency-bound: operand loads after the gathers, profiles/r02_forwarding_kernel_ncu_
// summary.json) every global operand of the stage is requested BEFORE the wait on the tile copy and folded into the
// accumulators as it arrives, so a CTA has one exposed memory latency.
template <bool REAL_G, int TBITS, int RB>
__global__ void __launch_bounds__(1 << (TBITS - RB), 2) stage_d2_fwd_kernel(const __grid_constant__ StageArgs2 m) {
    constexpr int R = 1 << RB;
    constexpr int NT = 1 << (TBITS - RB);
    constexpr int TSIZE = 1 << TBITS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c2* tile = reinterpret_cast<c2*>(smem_raw);
    c2* rtile = tile + TSIZE;
    __shared__ __align__(8) uint64_t mbar;

    const int chain = blockIdx.y / m.n_traj;
    const StageArgs& a = m.a[chain];
    const PassGeom g = a.geo;
    const int tid = threadIdx.x;
    const long long traj = blockIdx.y - chain * m.n_traj;
    const long long base = tile_base_of(g, blockIdx.x);
    const long long voff = traj * a.D;
    const c2* vsrc = a.v + voff;

    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    pdl_wait();
    pdl_launch_dependents();
    if (tid == 0) mbar_arrive_expect_tx(&mbar, (uint32_t)TSIZE * 16u);
    {
        const int rows = 1 << g.hi_bits;
        const uint32_t row_bytes = (uint32_t)(16u << g.lo_bits);
        for (int r = tid; r < rows; r += NT)
            tma_load_1d(tile + ((size_t)r << g.lo_bits), vsrc + base + ((long long)r << g.hi_shift), row_bytes, &mbar);
    }
    const long long lomask = (1LL << g.lo_bits) - 1;
    const long long fixed = base | (tid & lomask) | ((long long)(tid >> g.lo_bits) << g.hi_shift);
    // the register-block bits are the top RB tile bits: their global positions
    int pq[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) {
        const int j = TBITS - RB + q;
        pq[q] = (j < g.lo_bits) ? j : (j - g.lo_bits + g.hi_shift);
    }
    auto idx_of = [&](int r) {
        long long o = fixed;
#pragma unroll
        for (int q = 0; q < RB; ++q) o |= (long long)((r >> q) & 1) << pq[q];
        return o;
    };
    // ---- every global operand goes out now; each is folded into an accumulator as soon as it is used ----
    double pr[R], pi[R], qr[R], qi[R];
    c2 part[R];      // c_psi psi + c_b2 b2
    double dg[R];    // scaled diagonal of the amplitude
    {
        const double* dsrc = a.dint ? a.dint + traj * a.dint_stride : nullptr;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long ix = idx_of(r);
            c2 w0 = {0.0, 0.0}, w1 = {0.0, 0.0}, ps = {0.0, 0.0}, bb = {0.0, 0.0};
            if (a.w_in) { w0 = ld_own(a.w_in + voff + ix); if (!REAL_G) w1 = ld_own(a.w_in + a.w_plane + voff + ix); }
            if (a.psi) ps = ld_own(a.psi + voff + ix);
            if (a.b2) bb = ld_own(a.b2 + voff + ix);
            const double dv = dsrc ? __ldcs(dsrc + ix) : 0.0;
            pr[r] = w0.x; pi[r] = w0.y; qr[r] = w1.x; qi[r] = w1.y;
            part[r] = cadd(cmul(a.coef.c_psi, ps), cmul(a.coef.c_b2, bb));
            const int ones = __popcll((unsigned long long)ix);
            const int cnt = a.from_is_one ? ones : (g.n_bits - ones);
            dg[r] = fma(-a.u.theta, (double)cnt, fma(a.u.w, dv, -a.u.gamma));
        }
    }
    const int to_bit = a.to_bit;
    // partners across the bits above both geometries (N > TBITS + hb): coalesced loads, also ahead of the wait
    for (unsigned long long em = g.extra_mask; em; em &= em - 1) {
        const int p = __ffsll((long long)em) - 1;
        const double sg = (((base >> p) & 1) == to_bit) ? 1.0 : -1.0;
        double2 raw[R];
#pragma unroll
        for (int r = 0; r < R; ++r) raw[r] = __ldg(reinterpret_cast<const double2*>(vsrc + (idx_of(r) ^ (1LL << p))));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            pr[r] += raw[r].x; pi[r] += raw[r].y;
            if (!REAL_G) { qr[r] = fma(sg, raw[r].x, qr[r]); qi[r] = fma(sg, raw[r].y, qi[r]); }
        }
    }
    mbar_wait(&mbar, 0);
    const int jstart = __ffs(g.tile_flip_mask) - 1;
    c2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = tile[tid + r * NT];
    rb_tile_gather<true, REAL_G, TBITS, RB>(g, tile, nullptr, tid, to_bit, jstart, false, v, pr, pi, qr, qi);
    // ---- epilogue: out = part + c_g (diag v + g (P + w_in)) ----
#pragma unroll
    for (int r = 0; r < R; ++r) {
        double dx = a.u.g.x * pr[r], dy = a.u.g.x * pi[r];
        if (!REAL_G) { dx = fma(-a.u.g.y, qi[r], dx); dy = fma(a.u.g.y, qr[r], dy); }
        const c2 gv = {fma(dg[r], v[r].x, dx), fma(dg[r], v[r].y, dy)};
        const c2 res = cadd(part[r], cmul(a.coef.c_g, gv));
        st_c2(a.out + voff + idx_of(r), res);
        v[r] = res;
    }
    if (a.w_out) {
        // sums of the RESULT over this tile's flips for the next stage (whose tile is not closed under them)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            rtile[tid + r * NT] = v[r];
            pr[r] = 0.0; pi[r] = 0.0; qr[r] = 0.0; qi[r] = 0.0;
        }
        __syncthreads();
        rb_tile_gather<true, REAL_G, TBITS, RB>(g, rtile, nullptr, tid, to_bit, jstart, false, v, pr, pi, qr, qi);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long ix = idx_of(r);
            st_c2(a.w_out + voff + ix, c2{pr[r], pi[r]});
            if (!REAL_G) st_c2(a.w_out + a.w_plane + voff + ix, c2{qr[r], qi[r]});
        }
    }
}

// ---- d = 2 stage kernel of the time-dependent Taylor propagator (one drive time shape of constant phase) ---------
// On a step [a, a+h] the interpolated coefficients (QobjEvo's cubic splines, hamiltonian.py:436) are polynomials in
// u = (t-a)/h:  H(u) = sum_j H_j u^j,  H_0 = Dint - th_0 n_from - gam_0 + om_0 X,  H_j = -th_j n_from - gam_j + om_j X
// with X = sum_k (unit |to><from|_k + h.c.).  psi(u) = sum_k chi_k u^k solves psi' = -i h H(u) psi exactly when
//     (k+1) chi_{k+1} = -i h sum_{j <= min(p,k)} H_j chi_{k-j} ,
// i.e. ONE gather G_k = X chi_k per order and own-element history terms: no Magnus commutator error, no inner
// products, no host synchronisation; the step length is bounded by the spectral width (rho = h W, fp64
// cancellation: rho <= 14) and by the polynomial fit of the splines only.  The stage computes chi_{k+1} from the tile of chi_k,
// optionally stores G_k for later orders and folds chi_k + chi_{k+1} into the accumulator of psi(1) on every other
// order.  Replaces qutip.sesolve (simulation.py:729-735) for global drives of constant phase, and -- with per-qubit
// static factors from a per-trajectory table (TaylorArgs::table) -- the trajectory loop's solves (simulation.py:885-915).
#define PB200_TAYLOR_PMAX 8
struct TaylorArgs {
    const c2* v;       // chi_k, gather source [B][D]
    c2* out;           // chi_{k+1}
    c2* g_out;         // G_k = X chi_k (nullptr: nobody reads it later)
    c2* acc;           // accumulator of sum_k chi_k
    const double* dint;  // nullptr: no interaction
    long long dint_stride;  // 0: Dint shared by the trajectories
    long long D;
    PassGeom geo;
    c2 unit;           // uniform drive: e^{-i phi}, the constant phase
    // separable per-qubit drives (trajectory batches with static noise): coef_{b,k}(t) = a_{b,k} unit omega(t),
    // det_{b,k}(t) = theta(t) + c_{b,k} M(t).  table[b][0 .. 2N) = a unit per BIT position (re, im),
    // table[b][2N .. 3N) = c per bit position (layout of d2_table_stride); nullptr in the uniform case
    const double* table;
    int to_bit, from_is_one;
    double th0, gam0, om0, m0;  // H_0 = Dint - th0 n_from - m0 sum_k c_k n_k - gam0 + om0 X
    c2 scale;          // -i h / (k+1)
    int nh;            // history terms j = 1 .. nh
    const c2* hchi[PB200_TAYLOR_PMAX];   // chi_{k-j}   (nullptr when th_j = m_j = gam_j = 0)
    const c2* hg[PB200_TAYLOR_PMAX];     // G_{k-j}     (nullptr when om_j = 0)
    double hth[PB200_TAYLOR_PMAX], hgam[PB200_TAYLOR_PMAX], hom[PB200_TAYLOR_PMAX], hm[PB200_TAYLOR_PMAX];
    int acc_read;      // 1: acc is read before it is updated (0: first write of the step)
    int acc_add_v;     // 1: chi_k joins the update (even orders), 0: chi_{k+1} alone
    int acc_on;        // 0: this order leaves the accumulator alone
    c2 acc_mul;        // factor of the whole accumulator (phase of the scalar centre on the last order, else 1)
};

// epilogue of one amplitude block: everything after the partner sums.  `off[r]` = sum_k c_k [digit_k == from] of the
// amplitude (0 for uniform drives); idx is the index inside the trajectory, voff the trajectory's offset.
template <int R>
__device__ __forceinline__ void taylor_epilogue(const TaylorArgs& a, const long long (&idx)[R], const c2 (&v)[R],
                                                const double (&gx)[R], const double (&gy)[R], const double (&off)[R],
                                                long long voff, const double* __restrict__ dsrc) {
    constexpr int H = (R >= 4) ? R / 2 : R;
    const int nb = a.geo.n_bits;
#pragma unroll
    for (int h0 = 0; h0 < R; h0 += H) {
        double sx[H], sy[H], cn[H];
        {
            double dv[H];
#pragma unroll
            for (int r = 0; r < H; ++r) dv[r] = dsrc ? __ldcs(dsrc + idx[h0 + r]) : 0.0;
#pragma unroll
            for (int r = 0; r < H; ++r) {
                const int ones = __popcll((unsigned long long)idx[h0 + r]);
                cn[r] = (double)(a.from_is_one ? ones : (nb - ones));
                const double diag = fma(-a.th0, cn[r], fma(-a.m0, off[h0 + r], dv[r] - a.gam0));
                sx[r] = fma(diag, v[h0 + r].x, a.om0 * gx[h0 + r]);
                sy[r] = fma(diag, v[h0 + r].y, a.om0 * gy[h0 + r]);
            }
        }
        for (int j = 0; j < a.nh; ++j) {
            if (a.hchi[j]) {
                c2 c[H];
#pragma unroll
                for (int r = 0; r < H; ++r) c[r] = ld_own(a.hchi[j] + voff + idx[h0 + r]);
#pragma unroll
                for (int r = 0; r < H; ++r) {
                    const double d = -fma(a.hth[j], cn[r], fma(a.hm[j], off[h0 + r], a.hgam[j]));
                    sx[r] = fma(d, c[r].x, sx[r]); sy[r] = fma(d, c[r].y, sy[r]);
                }
            }
            if (a.hg[j]) {
                c2 c[H];
#pragma unroll
                for (int r = 0; r < H; ++r) c[r] = ld_own(a.hg[j] + voff + idx[h0 + r]);
#pragma unroll
                for (int r = 0; r < H; ++r) { sx[r] = fma(a.hom[j], c[r].x, sx[r]); sy[r] = fma(a.hom[j], c[r].y, sy[r]); }
            }
        }
        c2 res[H];
#pragma unroll
        for (int r = 0; r < H; ++r) {
            res[r] = {a.scale.x * sx[r] - a.scale.y * sy[r], a.scale.x * sy[r] + a.scale.y * sx[r]};
            st_c2(a.out + voff + idx[h0 + r], res[r]);
            if (a.g_out) st_c2(a.g_out + voff + idx[h0 + r], c2{gx[h0 + r], gy[h0 + r]});
        }
        if (a.acc_on) {
            c2 ac[H];
#pragma unroll
            for (int r = 0; r < H; ++r) ac[r] = a.acc_read ? ld_own(a.acc + voff + idx[h0 + r]) : c2{0.0, 0.0};
#pragma unroll
            for (int r = 0; r < H; ++r) {
                c2 s = cadd(ac[r], res[r]);
                if (a.acc_add_v) s = cadd(s, v[h0 + r]);
                st_c2(a.acc + voff + idx[h0 + r], cmul(a.acc_mul, s));
            }
        }
    }
}

// UNIFORM: one drive coefficient for every qubit and a single state (C2, C5); otherwise per-(trajectory, qubit) static
// factors from `table`, blockIdx.y = trajectory (C4: doppler + amplitude noise batches).
template <bool UNIFORM, bool REAL_G, int TBITS, int RB>
__global__ void __launch_bounds__(1 << (TBITS - RB), (65536 / ((1 << (TBITS - RB)) * (RB >= 3 ? 128 : 64))))
stage_d2_taylor_kernel(const __grid_constant__ TaylorArgs a) {
    constexpr int R = 1 << RB;
    constexpr int NT = 1 << (TBITS - RB);
    constexpr int TSIZE = 1 << TBITS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c2* tile = reinterpret_cast<c2*>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    const PassGeom& g = a.geo;
    const int tid = threadIdx.x;
    const long long traj = UNIFORM ? 0 : (long long)blockIdx.y;
    const long long voff = traj * a.D;
    const long long base = tile_base_of(g, blockIdx.x);
    const c2* vsrc = a.v + voff;

    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    pdl_wait();
    pdl_launch_dependents();
    if (tid == 0) mbar_arrive_expect_tx(&mbar, (uint32_t)TSIZE * 16u);
    {
        const int rows = 1 << g.hi_bits;
        const uint32_t row_bytes = (uint32_t)(16u << g.lo_bits);
        for (int r = tid; r < rows; r += NT)
            tma_load_1d(tile + ((size_t)r << g.lo_bits), vsrc + base + ((long long)r << g.hi_shift), row_bytes, &mbar);
    }
    double* tab = reinterpret_cast<double*>(tile + TSIZE);
    if (!UNIFORM) {   // the per-trajectory table is read behind the tile copy
        const int stride = d2_table_stride(g.n_bits);
        const double* src = a.table + traj * stride;
        for (int i = tid; i < stride; i += NT) tab[i] = src[i];
        __syncthreads();
    }
    const long long lomask = (1LL << g.lo_bits) - 1;
    const int to_bit = a.to_bit;
    c2 v[R];
    double pr[R], pi[R], qr[R], qi[R];
    long long idx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int t = tid + r * NT;
        idx[r] = base | (t & lomask) | ((long long)(t >> g.lo_bits) << g.hi_shift);
        pr[r] = 0.0; pi[r] = 0.0; qr[r] = 0.0; qi[r] = 0.0;
    }
    // partners across the bits outside the tile: coalesced loads issued while the bulk copy of the tile is in flight
    for (unsigned long long m = g.extra_mask; m; m &= m - 1) {
        const int p = __ffsll((long long)m) - 1;
        const int bit = (int)((base >> p) & 1);
        const double sg = (bit == to_bit) ? 1.0 : -1.0;
        double gx = 0.0, gy = 0.0;
        if (!UNIFORM) { gx = tab[2 * p]; gy = (bit == to_bit) ? tab[2 * p + 1] : -tab[2 * p + 1]; }
        double2 raw[R];
#pragma unroll
        for (int r = 0; r < R; ++r) raw[r] = __ldg(reinterpret_cast<const double2*>(vsrc + (idx[r] ^ (1LL << p))));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (UNIFORM) {
                pr[r] += raw[r].x; pi[r] += raw[r].y;
                if (!REAL_G) { qr[r] = fma(sg, raw[r].x, qr[r]); qi[r] = fma(sg, raw[r].y, qi[r]); }
            } else {
                pr[r] = fma(gx, raw[r].x, pr[r]); pr[r] = fma(-gy, raw[r].y, pr[r]);
                pi[r] = fma(gx, raw[r].y, pi[r]); pi[r] = fma(gy, raw[r].x, pi[r]);
            }
        }
    }
    mbar_wait(&mbar, 0);
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = tile[tid + r * NT];
    rb_tile_gather<UNIFORM, REAL_G, TBITS, RB>(g, tile, tab, tid, to_bit, 0, false, v, pr, pi, qr, qi);
    double off[R];
    if (UNIFORM) {
        // G = unit S_to + conj(unit) S_from = ux P + i uy Q
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double dx = a.unit.x * pr[r], dy = a.unit.x * pi[r];
            if (!REAL_G) { pr[r] = fma(-a.unit.y, qi[r], dx); pi[r] = fma(a.unit.y, qr[r], dy); }
            else { pr[r] = dx; pi[r] = dy; }
            off[r] = 0.0;
        }
    } else {
        // static per-qubit detuning weights: sum over (bits of base) + (bits of tid) + (register bits)
        const int nb = g.n_bits;
        const long long fixed = base | (tid & lomask) | ((long long)(tid >> g.lo_bits) << g.hi_shift);
        double common = 0.0;
        for (int p = 0; p < nb; ++p) {
            const int bit = (int)((fixed >> p) & 1);
            common += (bit == a.from_is_one) ? tab[2 * nb + p] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double acc = common;
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                const int j = TBITS - RB + q;
                const int p = (j < g.lo_bits) ? j : (j - g.lo_bits + g.hi_shift);
                const double th = tab[2 * nb + p];
                const int bit = (r >> q) & 1;   // `fixed` has these bits at 0
                acc += ((bit == a.from_is_one) ? th : 0.0) - ((0 == a.from_is_one) ? th : 0.0);
            }
            off[r] = acc;
        }
    }
    taylor_epilogue<R>(a, idx, v, pr, pi, off, voff, a.dint ? a.dint + traj * a.dint_stride : nullptr);
}

// any register size (N < 11 in particular): one thread per amplitude, partners through global loads
__global__ void __launch_bounds__(256) stage_d2_taylor_small_kernel(const __grid_constant__ TaylorArgs a) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.D) return;
    const int nb = a.geo.n_bits;
    const long long traj = blockIdx.y;
    const long long voff = traj * a.D;
    const double* tab = a.table ? a.table + traj * d2_table_stride(nb) : nullptr;
    double gxs = 0.0, gys = 0.0, offv = 0.0;
    if (tab) {
        for (int p = 0; p < nb; ++p) {
            const double2 raw = __ldg(reinterpret_cast<const double2*>(a.v + voff + (s ^ (1LL << p))));
            const int bit = (int)((s >> p) & 1);
            const double gx = tab[2 * p], gy = (bit == a.to_bit) ? tab[2 * p + 1] : -tab[2 * p + 1];
            gxs = fma(gx, raw.x, gxs); gxs = fma(-gy, raw.y, gxs);
            gys = fma(gx, raw.y, gys); gys = fma(gy, raw.x, gys);
            offv += (bit == a.from_is_one) ? tab[2 * nb + p] : 0.0;
        }
    } else {
        double pr = 0.0, pi = 0.0, qr = 0.0, qi = 0.0;
        for (int p = 0; p < nb; ++p) {
            const double2 raw = __ldg(reinterpret_cast<const double2*>(a.v + voff + (s ^ (1LL << p))));
            const double sg = ((int)((s >> p) & 1) == a.to_bit) ? 1.0 : -1.0;
            pr += raw.x; pi += raw.y; qr = fma(sg, raw.x, qr); qi = fma(sg, raw.y, qi);
        }
        gxs = fma(-a.unit.y, qi, a.unit.x * pr);
        gys = fma(a.unit.y, qr, a.unit.x * pi);
    }
    const long long idx[1] = {s};
    const double2 own = __ldg(reinterpret_cast<const double2*>(a.v + voff + s));
    const c2 v[1] = {{own.x, own.y}};
    const double gx[1] = {gxs}, gy[1] = {gys}, off[1] = {offv};
    taylor_epilogue<1>(a, idx, v, gx, gy, off, voff, a.dint ? a.dint + traj * a.dint_stride : nullptr);
}

// ---- generic-d stage kernel (any dim, several drives; global gathers) -------
// table per (exponential, trajectory):
//   for each drive q: g[q][k] (re,im) per QUDIT k, theta[q][k]; then w, gamma
// per-(exponential, trajectory) table of the generic kernel: per drive q [g (re,im) per qudit | theta per qudit],
// then wc (weight of the SLM-masked part of the interaction, XY mode), w, gamma
__host__ __device__ inline int gen_table_stride(int n, int n_drives) { return n_drives * 3 * n + 3; }

#define PB200_TILED_MAX_HIGH 40
#define PB200_MAX_DRIVES_K 3
struct GenArgs {
    const c2* v; const c2* psi; const c2* b2; c2* out;
    const double* dint; long long dint_stride; long long D;
    int n, dim, n_drives;
    int to[3], from[3];
    StageCoef coef;
    const double* table;  // [B][stride]
    const double* beta_dev;
    // XY mode: exchange couplings U^xy_ij (|u d><d u| + h.c.), [Bx][n*n], weighted by the table's w like Dint
    const double* xy; long long xy_stride; int xy_u, xy_d;
    // XY mode with an SLM mask (hamiltonian.py:399-424): pairs touching a masked qudit (bit k of slm_mask) carry the
    // weight wc = table[stride - 3] instead of w; dint2 = interaction diagonal of those pairs (dint: the others)
    unsigned long long slm_mask; const double* dint2;
    double* dot_acc;   // fused reductions (register-blocked tiled kernel only): acc[traj][0] += Re<lhs, out>, [1] += <out, out>
    LanczosFuse lz;    // fused Lanczos step when lz.vj != nullptr (register-blocked tiled kernel only)
};

__global__ void __launch_bounds__(256) stage_generic_kernel(GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);
    const int stride = gen_table_stride(a.n, a.n_drives);
    const long long traj = blockIdx.y;
    for (int i = threadIdx.x; i < stride; i += blockDim.x) tab[i] = a.table[traj * stride + i];
    __syncthreads();
    double* xys = tab + stride;  // XY couplings of this trajectory, weighted like Dint
    if (a.xy) {
        const double wx = tab[stride - 2], wxc = tab[stride - 3];
        for (int i = threadIdx.x; i < a.n * a.n; i += blockDim.x) {
            const int qi = i / a.n, qj = i - qi * a.n;
            const bool touched = ((a.slm_mask >> qi) | (a.slm_mask >> qj)) & 1ULL;
            xys[i] = a.xy[traj * a.xy_stride + i] * (touched ? wxc : wx);
        }
        __syncthreads();
    }
    const double w = tab[stride - 2], gamma = tab[stride - 1];
    const long long voff = traj * a.D;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < a.D;
         idx += (long long)gridDim.x * blockDim.x) {
        const c2 vo = a.v[voff + idx];
        double diag = -gamma;
        if (a.dint) diag = fma(w, a.dint[traj * a.dint_stride + idx], diag);
        if (a.dint2) diag = fma(tab[stride - 3], a.dint2[traj * a.dint_stride + idx], diag);
        double rr = 0.0, ri = 0.0;
        long long rem = idx, st = 1;
        for (int k = a.n - 1; k >= 0; --k) {  // qudit k has stride dim^(n-1-k)
            const int digit = (int)(rem % a.dim);
            rem /= a.dim;
            for (int q = 0; q < a.n_drives; ++q) {
                const double* gq = tab + q * 3 * a.n;
                if (digit == a.to[q]) {
                    const c2 pv = a.v[voff + idx + (long long)(a.from[q] - a.to[q]) * st];
                    const double gx = gq[2 * k], gy = gq[2 * k + 1];
                    rr = fma(gx, pv.x, rr); rr = fma(-gy, pv.y, rr);
                    ri = fma(gx, pv.y, ri); ri = fma(gy, pv.x, ri);
                } else if (digit == a.from[q]) {
                    const c2 pv = a.v[voff + idx + (long long)(a.to[q] - a.from[q]) * st];
                    const double gx = gq[2 * k], gy = -gq[2 * k + 1];
                    rr = fma(gx, pv.x, rr); rr = fma(-gy, pv.y, rr);
                    ri = fma(gx, pv.y, ri); ri = fma(gy, pv.x, ri);
                    diag -= gq[2 * a.n + k];
                }
            }
            st *= a.dim;
        }
        if (a.xy) {  // flip-flop partners: every pair (i, j) holding (u, d) or (d, u)  (make_xy_term, :276-294)
            signed char dg[40];
            long long sts[40];
            long long r2 = idx, s2 = 1;
            for (int k = a.n - 1; k >= 0; --k) { dg[k] = (signed char)(r2 % a.dim); r2 /= a.dim; sts[k] = s2; s2 *= a.dim; }
            for (int i = 0; i < a.n; ++i) {
                if (dg[i] != a.xy_u && dg[i] != a.xy_d) continue;
                for (int j = i + 1; j < a.n; ++j) {
                    if ((dg[i] == a.xy_u && dg[j] == a.xy_d) || (dg[i] == a.xy_d && dg[j] == a.xy_u)) {
                        const double u = xys[i * a.n + j];
                        if (u == 0.0) continue;
                        const long long pidx2 = idx + (long long)(dg[j] - dg[i]) * sts[i] + (long long)(dg[i] - dg[j]) * sts[j];
                        const c2 pv = a.v[voff + pidx2];
                        rr = fma(u, pv.x, rr); ri = fma(u, pv.y, ri);
                    }
                }
            }
        }
        c2 gv = {fma(diag, vo.x, rr), fma(diag, vo.y, ri)};
        c2 res = cmul(a.coef.c_g, gv);
        if (a.psi) res = cadd(res, cmul(a.coef.c_psi, a.psi[voff + idx]));
        if (a.b2) res = cadd(res, cmul(a.beta_dev ? c2{-a.beta_dev[traj], 0.0} : a.coef.c_b2, a.b2[voff + idx]));
        st_c2(a.out + voff + idx, res);
    }
}

// ---- tiled stage kernel for d = 3 / 4 (the "all" basis, leakage levels) ---------------------------------------
// Same maths as stage_generic_kernel without the XY exchange term.  A CTA owns the DIM^K amplitudes that share
// their n - K most significant digits (one contiguous run, brought in by ONE TMA bulk copy): partners across the
// K low digits are shared-memory reads selected arithmetically (no divergence: a digit that is neither |to> nor
// |from> of a drive reads itself with a zero coefficient); the high digits are the same for the whole tile, so
// their partners are a short CTA-uniform list of (offset, coefficient) pairs served by coalesced loads.
struct TileExtra { long long off; double gx, gy; };

// ---- register-blocked tiled stage kernel for d = 3 / 4 --------------------------------------------------------
// stage_tiled_kernel is instruction-bound (ncu on C3: 1053 thread instructions per amplitude, issue-active 62 %,
// DRAM 10 %: the digit decomposition, the coefficient selects and the table reads are redone for every amplitude).
// Here a thread owns the R = DIM^RBD amplitudes that differ in the top RBD tile digits and keeps their
// accumulators in registers: digits, selects and coefficients are computed once per (digit, drive) and reused for
// the R amplitudes, exactly as the d = 2 kernel reuses them across its register block.
__host__ __device__ constexpr int ipow_c(int b, int e) { return e <= 0 ? 1 : b * ipow_c(b, e - 1); }

template <int DIM, int K, int RBD>
__global__ void __launch_bounds__(256, 2) stage_multilevel_rb_kernel(GenArgs a) {
    constexpr int R = ipow_c(DIM, RBD);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ TileExtra extra[PB200_MAX_DRIVES_K * PB200_TILED_MAX_HIGH];
    __shared__ int n_extra;
    __shared__ double diag_high;

    const int tid = threadIdx.x;
    const long long traj = blockIdx.y;
    const int n = a.n;
    const int kk = n < K ? n : K;           // digits inside the tile (host guarantees kk >= RBD)
    int nt_act = 1;
    for (int j = 0; j < kk - RBD; ++j) nt_act *= DIM;   // active threads = stride of the first register digit
    const int tsz = nt_act * R;
    c2* tile = reinterpret_cast<c2*>(smem_raw);
    double* tab = reinterpret_cast<double*>(smem_raw + (((size_t)tsz * 16 + 127) / 128) * 128);
    const int stride = gen_table_stride(n, a.n_drives);
    const long long base = (long long)blockIdx.x * tsz;
    const long long voff = traj * a.D;

    if (tid == 0) mbar_init(&mbar, 1);
    for (int i = tid; i < stride; i += blockDim.x) tab[i] = a.table[traj * stride + i];
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&mbar, (uint32_t)tsz * 16u);
        tma_load_1d(tile, a.v + voff + base, (uint32_t)tsz * 16u, &mbar);
    }
    if (tid < 32) {
        // partners across the digits above the tile: one lane per (digit, drive), compacted in order by ballot
        int cnt = 0;
        double dh = 0.0;
        const int total = (n - kk) * a.n_drives;
        for (int e0 = 0; e0 < total; e0 += 32) {
            const int e = e0 + tid;
            bool valid = false;
            TileExtra te = {0, 0.0, 0.0};
            if (e < total) {
                const int j = kk + e / a.n_drives, q = e % a.n_drives;
                long long rem = blockIdx.x, st = tsz;
                for (int jj = kk; jj < j; ++jj) { rem /= DIM; st *= DIM; }
                const int digit = (int)(rem % DIM);
                const int k = n - 1 - j;
                const double* gq = tab + q * 3 * n;
                if (digit == a.to[q]) {
                    valid = true;
                    te = {(long long)(a.from[q] - a.to[q]) * st, gq[2 * k], gq[2 * k + 1]};
                } else if (digit == a.from[q]) {
                    valid = true;
                    te = {(long long)(a.to[q] - a.from[q]) * st, gq[2 * k], -gq[2 * k + 1]};
                    dh -= gq[2 * n + k];
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            if (valid) extra[cnt + __popc(m & ((1u << tid) - 1u))] = te;
            cnt += __popc(m);
        }
        for (int o = 16; o > 0; o >>= 1) dh += __shfl_xor_sync(0xffffffffu, dh, o);
        if (tid == 0) { n_extra = cnt; diag_high = dh; }
    }
    __syncthreads();
    mbar_wait(&mbar, 0);
    double dot0 = 0.0, dot1 = 0.0;
    const bool fuse = a.lz.vj != nullptr;
    LanczosCoef lc{0.0, 0.0, 0.0, 0.0};
    if (fuse) lc = lanczos_coef(a.lz, traj);
    if (tid < nt_act) {

    double rr[R], ri[R], dd[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { rr[i] = 0.0; ri[i] = 0.0; dd[i] = 0.0; }
    double diag_low = 0.0;
    // --- the tile digits below the register block: selects once per (digit, drive), R shared-memory reads ---
    {
        int rem = tid, st = 1;
#pragma unroll
        for (int j = 0; j < K - RBD; ++j) {
            if (j < kk - RBD) {
                const int digit = rem % DIM;
                rem /= DIM;
                const int k = n - 1 - j;
                for (int q = 0; q < a.n_drives; ++q) {
                    const double* gq = tab + q * 3 * n;
                    const bool is_to = digit == a.to[q], is_from = digit == a.from[q];
                    if (is_to || is_from) {
                        const int off = tid + (is_to ? (a.from[q] - a.to[q]) : (a.to[q] - a.from[q])) * st;
                        const double gx = gq[2 * k];
                        const double gy = is_to ? gq[2 * k + 1] : -gq[2 * k + 1];
                        diag_low -= is_from ? gq[2 * n + k] : 0.0;
#pragma unroll
                        for (int i = 0; i < R; ++i) {
                            const c2 pv = tile[off + i * nt_act];
                            rr[i] = fma(gx, pv.x, rr[i]); rr[i] = fma(-gy, pv.y, rr[i]);
                            ri[i] = fma(gx, pv.y, ri[i]); ri[i] = fma(gy, pv.x, ri[i]);
                        }
                    }
                }
                st *= DIM;
            }
        }
    }
    // --- the register-block digits: the digit of amplitude i is a compile-time constant ---
#pragma unroll
    for (int jj = 0; jj < RBD; ++jj) {
        const int j = kk - RBD + jj;
        const int k = n - 1 - j;
        const int stj = nt_act * ipow_c(DIM, jj);
        for (int q = 0; q < a.n_drives; ++q) {
            const double* gq = tab + q * 3 * n;
            const double gx0 = gq[2 * k], gy0 = gq[2 * k + 1], th = gq[2 * n + k];
            const int to = a.to[q], from = a.from[q];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int digit = (i / ipow_c(DIM, jj)) % DIM;
                const bool is_to = digit == to, is_from = digit == from;
                if (is_to || is_from) {   // uniform across the CTA
                    const c2 pv = tile[tid + i * nt_act + (is_to ? (from - to) : (to - from)) * stj];
                    const double gy = is_to ? gy0 : -gy0;
                    rr[i] = fma(gx0, pv.x, rr[i]); rr[i] = fma(-gy, pv.y, rr[i]);
                    ri[i] = fma(gx0, pv.y, ri[i]); ri[i] = fma(gy, pv.x, ri[i]);
                    dd[i] -= is_from ? th : 0.0;
                }
            }
        }
    }
    // --- the digits above the tile: CTA-uniform (offset, coefficient) list, coalesced loads ---
    const c2* vbase = a.v + voff + base + tid;
    const int nex = n_extra;
    for (int e = 0; e < nex; ++e) {
        const long long off = extra[e].off;
        const double gx = extra[e].gx, gy = extra[e].gy;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const double2 raw = __ldg(reinterpret_cast<const double2*>(vbase + off + i * nt_act));
            rr[i] = fma(gx, raw.x, rr[i]); rr[i] = fma(-gy, raw.y, rr[i]);
            ri[i] = fma(gx, raw.y, ri[i]); ri[i] = fma(gy, raw.x, ri[i]);
        }
    }
    // --- epilogue ---
    const double w = tab[stride - 2], gamma = tab[stride - 1];
    const c2 cb2 = a.beta_dev ? c2{-a.beta_dev[traj], 0.0} : a.coef.c_b2;
    const double dcommon = diag_high + diag_low - gamma;
    const double* dsrc = a.dint ? a.dint + traj * a.dint_stride : nullptr;
    constexpr int H = (R % 3 == 0) ? 3 : 4;
#pragma unroll
    for (int h0 = 0; h0 < R; h0 += H) {
        double dv[H];
        c2 pv[H], bv[H];
#pragma unroll
        for (int r = 0; r < H; ++r) {
            const long long idx = base + tid + (long long)(h0 + r) * nt_act;
            dv[r] = dsrc ? __ldcs(dsrc + idx) : 0.0;
            pv[r] = {0.0, 0.0}; bv[r] = {0.0, 0.0};
            if (fuse) {
                pv[r] = ld_own(a.lz.vj + voff + idx);
                if (a.lz.vjm1 && lc.beta_prev != 0.0) bv[r] = ld_own(a.lz.vjm1 + voff + idx);
            } else {
                if (a.psi) pv[r] = ld_own(a.psi + voff + idx);
                if (a.b2) bv[r] = ld_own(a.b2 + voff + idx);
            }
        }
#pragma unroll
        for (int r = 0; r < H; ++r) {
            const int i = h0 + r;
            const long long idx = base + tid + (long long)i * nt_act;
            const c2 vo = tile[tid + i * nt_act];
            const double diag = fma(w, dv[r], dcommon + dd[i]);
            const c2 gv = {fma(diag, vo.x, rr[i]), fma(diag, vo.y, ri[i])};
            c2 res, lhs;
            if (fuse) {   // fused Lanczos step (see LanczosFuse)
                const c2 vn = {(vo.x - lc.alpha * pv[r].x) * lc.inv, (vo.y - lc.alpha * pv[r].y) * lc.inv};
                const double ai = lc.alpha * lc.inv, abi = ai * lc.beta_prev;
                res.x = fma(lc.inv, gv.x, -fma(ai, vo.x, fma(abi, bv[r].x, lc.beta * pv[r].x)));
                res.y = fma(lc.inv, gv.y, -fma(ai, vo.y, fma(abi, bv[r].y, lc.beta * pv[r].y)));
                st_c2(a.lz.vout + voff + idx, vn);
                lhs = vn;
            } else {
                res = cmul(a.coef.c_g, gv);
                res = cadd(res, cmul(a.coef.c_psi, pv[r]));
                res = cadd(res, cmul(cb2, bv[r]));
                lhs = vo;
            }
            dot0 = fma(lhs.x, res.x, dot0); dot0 = fma(lhs.y, res.y, dot0);
            dot1 = fma(res.x, res.x, dot1); dot1 = fma(res.y, res.y, dot1);
            st_c2(a.out + voff + idx, res);
        }
    }
    }  // active threads
    if (a.dot_acc) {  // fused Lanczos inner products: warp __shfl reduction, one atomic pair per CTA
        for (int o = 16; o > 0; o >>= 1) {
            dot0 += __shfl_xor_sync(0xffffffffu, dot0, o);
            dot1 += __shfl_xor_sync(0xffffffffu, dot1, o);
        }
        __shared__ double dred[2][8];
        if ((tid & 31) == 0) { dred[0][tid >> 5] = dot0; dred[1][tid >> 5] = dot1; }
        __syncthreads();
        if (tid == 0) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { s0 += dred[0][i]; s1 += dred[1][i]; }
            atomicAdd(a.dot_acc + 2 * traj, s0);
            atomicAdd(a.dot_acc + 2 * traj + 1, s1);
        }
    }
    if (fuse && blockIdx.x == 0 && tid == 0) {
        a.lz.alpha_out[traj] = lc.alpha;
        a.lz.beta_out[traj] = lc.beta;
        a.lz.acc_clear[2 * traj] = 0.0;
        a.lz.acc_clear[2 * traj + 1] = 0.0;
    }
}

// ---- interaction diagonal ---------------------------------------------------
// Dint[s] = sum_{i<j} U_ij [digit_i == r][digit_j == r]
// (make_vdw_term, hamiltonian.py:260-274, after the + dag doubling of 0.5*U)
__global__ void dint_kernel(double* dint, const double* U, int n, int dim, int rstate, long long D) {
    extern __shared__ double Us[];
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) Us[i] = U[i];
    __syncthreads();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < D;
         idx += (long long)gridDim.x * blockDim.x) {
        int pos[64];
        int cnt = 0;
        long long rem = idx;
        for (int k = n - 1; k >= 0; --k) {
            if ((int)(rem % dim) == rstate) pos[cnt++] = k;
            rem /= dim;
        }
        double acc = 0.0;
        // pos is descending in k; accumulate pairs in (i<j) order of the reference loop
        for (int a = cnt - 1; a >= 0; --a)
            for (int b = a - 1; b >= 0; --b) acc += Us[pos[a] * n + pos[b]];
        dint[idx] = acc;
    }
}

// min / max of Dint grouped by the number of |r> digits (spectral bounds)
__global__ void dint_bounds_kernel(const double* dint, int n, int dim, int rstate, long long D, double* mins,
                                   double* maxs) {
    // one thread per amplitude, atomics on (n+1) bins via ordered-int trick
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < D;
         idx += (long long)gridDim.x * blockDim.x) {
        int cnt = 0;
        long long rem = idx;
        for (int k = 0; k < n; ++k) { cnt += ((int)(rem % dim) == rstate); rem /= dim; }
        const double v = dint[idx];
        // Dint >= 0 is not guaranteed (negative C6 never occurs, but be safe): use CAS loops
        unsigned long long* pmin = reinterpret_cast<unsigned long long*>(mins + cnt);
        unsigned long long old = *pmin;
        while (__longlong_as_double((long long)old) > v) {
            unsigned long long assumed = old;
            old = atomicCAS(pmin, assumed, (unsigned long long)__double_as_longlong(v));
            if (old == assumed) break;
        }
        unsigned long long* pmax = reinterpret_cast<unsigned long long*>(maxs + cnt);
        old = *pmax;
        while (__longlong_as_double((long long)old) < v) {
            unsigned long long assumed = old;
            old = atomicCAS(pmax, assumed, (unsigned long long)__double_as_longlong(v));
            if (old == assumed) break;
        }
    }
}

// ---- measurement: bitstring weights, occupations, sampling ---------------------------------------------------
// weights[b(s)] += |psi_s|^2 with bit k of b = [digit_k(s) == one_digit], qudit 0 = most significant bit
// (QutipResult._weights, qutip_result.py:101-158: reversal for ground-rydberg and the 3/4-level
// marginalisation are both this rule)
__global__ void bitstring_weights_kernel(const c2* psi, double* weights, long long D, int n, int dim, int one_digit) {
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < D;
         s += (long long)gridDim.x * blockDim.x) {
        const c2 v = psi[s];
        const double p = v.x * v.x + v.y * v.y;
        long long rem = s, b = 0;
        for (int k = n - 1; k >= 0; --k) {  // qudit k <-> bit n-1-k
            if ((int)(rem % dim) == one_digit) b |= 1LL << (n - 1 - k);
            rem /= dim;
        }
        if (dim == 2) weights[b] = p;  // a permutation: no atomics needed
        else atomicAdd(weights + b, p);
    }
}

// occ[k] += sum_s |psi_s|^2 [digit_k(s) == digit]   (Occupation observable / <n_k>)
__global__ void occupation_kernel(const c2* psi, double* occ, long long D, int n, int dim, int digit) {
    extern __shared__ double socc[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) socc[i] = 0.0;
    __syncthreads();
    const long long traj = blockIdx.y;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < D;
         s += (long long)gridDim.x * blockDim.x) {
        const c2 v = psi[traj * D + s];
        const double p = v.x * v.x + v.y * v.y;
        if (p == 0.0) continue;
        long long rem = s;
        for (int k = n - 1; k >= 0; --k) {
            if ((int)(rem % dim) == digit) atomicAdd(&socc[k], p);
            rem /= dim;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(occ + traj * n + i, socc[i]);
}

// corr[traj][i*n+j] (i <= j) += sum_s |psi_s|^2 [digit_i(s) == digit][digit_j(s) == digit]
// (CorrelationMatrix observable <n_i n_j>; the diagonal is the occupation).  A block stages 2048 probabilities and
// their per-qudit match masks in shared memory; each warp then reduces a subset of the n(n+1)/2 pairs over them.
__global__ void __launch_bounds__(256) correlation_kernel(const c2* psi, double* corr, long long D, int n, int dim, int digit) {
    constexpr int CH = 2048;
    __shared__ double sp[CH];
    __shared__ unsigned long long sm[CH];
    const long long traj = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int npairs = n * (n + 1) / 2;
    for (long long base = blockIdx.x * (long long)CH; base < D; base += (long long)gridDim.x * CH) {
        for (int e = threadIdx.x; e < CH; e += blockDim.x) {
            const long long s = base + e;
            double p = 0.0;
            unsigned long long m = 0ull;
            if (s < D) {
                const c2 v = psi[traj * D + s];
                p = v.x * v.x + v.y * v.y;
                long long rem = s;
                for (int k = n - 1; k >= 0; --k) {
                    if ((int)(rem % dim) == digit) m |= 1ull << k;
                    rem /= dim;
                }
            }
            sp[e] = p; sm[e] = m;
        }
        __syncthreads();
        int i = 0, first = 0;  // pairs enumerated row by row: (0,0..n-1), (1,1..n-1), ...
        for (int pr = warp; pr < npairs; pr += nw) {
            while (pr - first >= n - i) { first += n - i; ++i; }
            const int j = i + (pr - first);
            const unsigned long long need = (1ull << i) | (1ull << j);
            double acc = 0.0;
            for (int e = lane; e < CH; e += 32)
                if ((sm[e] & need) == need) acc += sp[e];
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0 && acc != 0.0) atomicAdd(corr + traj * n * n + i * n + j, acc);
        }
        __syncthreads();
    }
}

// acc[traj] += <phi, psi_traj> (complex; phi shared by all trajectories)   (Fidelity observable / State.overlap)
__global__ void overlap_kernel(const c2* phi, const c2* psi, long long D, double* acc) {
    const long long traj = blockIdx.y;
    double re = 0.0, im = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        const c2 a = phi[i], b = psi[traj * D + i];
        re = fma(a.x, b.x, re); re = fma(a.y, b.y, re);
        im = fma(a.x, b.y, im); im = fma(-a.y, b.x, im);
    }
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    __shared__ double ws[2][8];
    if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = re; ws[1][threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s0 += ws[0][i]; s1 += ws[1][i]; }
        atomicAdd(acc + 2 * traj, s0);
        atomicAdd(acc + 2 * traj + 1, s1);
    }
}

// indices[i] = first j with cum[j] >= u[i] * total  (np.searchsorted(cumsum(w / sum w), rnd), side="left")
__global__ void search_sorted_kernel(const double* cum, long long M, const double* u, long long* idx, int n_shots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_shots) return;
    const double target = u[i] * cum[M - 1];
    long long lo = 0, hi = M;  // first index with cum >= target
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (cum[mid] < target) lo = mid + 1; else hi = mid;
    }
    idx[i] = lo < M ? lo : M - 1;
}

// ---- small utilities --------------------------------------------------------
__global__ void set_basis_kernel(c2* psi, long long D, long long index) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x)
        psi[i] = {i == index ? 1.0 : 0.0, 0.0};
}

__global__ void prob_kernel(const c2* psi, double* probs, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const c2 v = psi[i];
        probs[i] = v.x * v.x + v.y * v.y;
    }
}

// squared norm per trajectory: grid.y = trajectory, warp-shuffle + one atomic per block
__global__ void norm2_kernel(const c2* psi, long long D, double* out) {
    const long long traj = blockIdx.y;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        const c2 v = psi[traj * D + i];
        acc = fma(v.x, v.x, acc);
        acc = fma(v.y, v.y, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += ws[i];
        atomicAdd(out + traj, s);
    }
}

// ---- dissipator of the Lindblad equation on the vectorised density matrix ------------------------------
// rho is stored as the state of 2N qudits (row digits above column digits).  A single-qudit collapse
// operator couples only the (row digit, column digit) pair of its qudit, so exp(h*D) factorises into one
// d^2 x d^2 matrix per qudit, applied in place to the pair of digits at strides s_hi > s_lo.
struct PairOp {
    c2 m[81];  // row-major [d*d][d*d], d <= 3
};

__global__ void pair_op_kernel(c2* psi, long long D, int dim, long long s_hi, long long s_lo,
                               const __grid_constant__ PairOp op) {
    const int dd = dim * dim;
    const long long groups = D / dd;
    const long long traj = blockIdx.y;
    c2* base = psi + traj * D;
    const long long mid_span = s_hi / (s_lo * dim);
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < groups;
         gidx += (long long)gridDim.x * blockDim.x) {
        long long q = gidx;
        const long long low = q % s_lo; q /= s_lo;
        const long long mid = q % mid_span; q /= mid_span;
        const long long idx0 = low + mid * s_lo * dim + q * s_hi * dim;
        c2 v[9], w[9];
        for (int a = 0; a < dim; ++a)
            for (int b = 0; b < dim; ++b) v[a * dim + b] = base[idx0 + a * s_hi + b * s_lo];
        for (int r = 0; r < dd; ++r) {
            double xr = 0.0, xi = 0.0;
            for (int c = 0; c < dd; ++c) {
                const c2 mm = op.m[r * dd + c];
                xr = fma(mm.x, v[c].x, xr); xr = fma(-mm.y, v[c].y, xr);
                xi = fma(mm.x, v[c].y, xi); xi = fma(mm.y, v[c].x, xi);
            }
            w[r] = {xr, xi};
        }
        for (int a = 0; a < dim; ++a)
            for (int b = 0; b < dim; ++b) base[idx0 + a * s_hi + b * s_lo] = w[a * dim + b];
    }
}

// ---- Lanczos (Krylov) propagator helpers ------------------------------------------------------------------
// acc[traj][0] += Re<v, w>, acc[traj][1] += <w, w>   (warp __shfl reduction, one atomic pair per block)
__global__ void dot2_kernel(const c2* v, const c2* w, long long D, double* acc) {
    const long long traj = blockIdx.y;
    double a0 = 0.0, a1 = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        const c2 x = v[traj * D + i], y = w[traj * D + i];
        a0 = fma(x.x, y.x, a0); a0 = fma(x.y, y.y, a0);
        a1 = fma(y.x, y.x, a1); a1 = fma(y.y, y.y, a1);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    __shared__ double ws[2][8];
    if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = a0; ws[1][threadIdx.x >> 5] = a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s0 += ws[0][i]; s1 += ws[1][i]; }
        atomicAdd(acc + 2 * traj, s0);
        atomicAdd(acc + 2 * traj + 1, s1);
    }
}

// w <- (w - alpha v) / beta with alpha = acc[0], beta = sqrt(acc[1] - alpha^2); records alpha, beta
// (beta = 0 and w = 0 on breakdown) and clears the accumulator of the other parity for the next iteration.
__global__ void lanczos_update_kernel(c2* w, const c2* v, long long D, const double* acc, double* alpha_out,
                                      double* beta_out, double* acc_clear) {
    const long long traj = blockIdx.y;
    const double alpha = acc[2 * traj];
    const double ww = acc[2 * traj + 1];
    const double b2 = ww - alpha * alpha;
    const double beta = (b2 > 1e-28 * fmax(ww, 1e-300)) ? sqrt(b2) : 0.0;
    const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        const c2 x = v[traj * D + i];
        c2 y = w[traj * D + i];
        y.x = (y.x - alpha * x.x) * inv;
        y.y = (y.y - alpha * x.y) * inv;
        w[traj * D + i] = y;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        alpha_out[traj] = alpha;
        beta_out[traj] = beta;
        acc_clear[2 * traj] = 0.0;
        acc_clear[2 * traj + 1] = 0.0;
    }
}

// v0 <- psi / ||psi|| with ||psi||^2 = acc[1]; records the norm
__global__ void normalize_copy_kernel(c2* v0, const c2* psi, long long D, const double* acc, double* norm_out,
                                      double* acc_clear) {
    const long long traj = blockIdx.y;
    const double nrm = sqrt(acc[2 * traj + 1]);
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        const c2 x = psi[traj * D + i];
        v0[traj * D + i] = {x.x * inv, x.y * inv};
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        norm_out[traj] = nrm;
        acc_clear[2 * traj] = 0.0;
        acc_clear[2 * traj + 1] = 0.0;
    }
}

// out = sum_j y[traj][j] V_j   (V_j = V + j * vstride; y interleaved complex [traj][m])
__global__ void krylov_combine_kernel(c2* out, const c2* V, long long vstride, long long D, const double* y, int m) {
    const long long traj = blockIdx.y;
    const double* yt = y + 2 * (long long)m * traj;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        double xr = 0.0, xi = 0.0;
        for (int j = 0; j < m; ++j) {
            const c2 v = V[j * vstride + traj * D + i];
            const double yr = yt[2 * j], yi = yt[2 * j + 1];
            xr = fma(yr, v.x, xr); xr = fma(-yi, v.y, xr);
            xi = fma(yr, v.y, xi); xi = fma(yi, v.x, xi);
        }
        out[traj * D + i] = {xr, xi};
    }
}

// ---- Monte-Carlo wave function: non-Hermitian decay and quantum jumps ---------------------------------------------
// psi[s] *= exp(-h/2 * sum_k gamma[digit_k(s)])  -- the diagonal part -i/2 sum L^+L of H_eff
struct DecayTable { double gamma[4]; };
__global__ void mcwf_decay_kernel(c2* psi, long long D, int n, int dim, double half_h, const __grid_constant__ DecayTable tb) {
    const long long traj = blockIdx.y;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < D;
         s += (long long)gridDim.x * blockDim.x) {
        long long rem = s;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) { acc += tb.gamma[(int)(rem % dim)]; rem /= dim; }
        const double f = exp(-half_h * acc);
        c2 v = psi[traj * D + s];
        v.x *= f; v.y *= f;
        psi[traj * D + s] = v;
    }
}

// psi <- scale * (L on qudit with stride `st`) psi for one trajectory; L row-major d x d
struct QuditOp { c2 m[16]; };
__global__ void qudit_op_kernel(c2* psi, long long D, int dim, long long st, double scale, const __grid_constant__ QuditOp op) {
    psi += (long long)blockIdx.y * D;   // gridDim.y = trajectories (1 for a single-trajectory jump)
    const long long groups = D / dim;
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < groups;
         gidx += (long long)gridDim.x * blockDim.x) {
        const long long low = gidx % st, high = gidx / st;
        const long long idx0 = low + high * st * dim;
        c2 v[4], w[4];
        for (int a = 0; a < dim; ++a) v[a] = psi[idx0 + a * st];
        for (int r = 0; r < dim; ++r) {
            double xr = 0.0, xi = 0.0;
            for (int c = 0; c < dim; ++c) {
                const c2 mm = op.m[r * dim + c];
                xr = fma(mm.x, v[c].x, xr); xr = fma(-mm.y, v[c].y, xr);
                xi = fma(mm.x, v[c].y, xi); xi = fma(mm.y, v[c].x, xi);
            }
            w[r] = {xr * scale, xi * scale};
        }
        for (int a = 0; a < dim; ++a) psi[idx0 + a * st] = w[a];
    }
}

// Single-qudit reduced density matrix rho[a][b] = sum_rest psi(a, rest) conj(psi(b, rest)) of the qudit with stride
// `st` (one trajectory), accumulated into acc[2 * (a * dim + b) + {0, 1}]: the jump weights <L^+L> of a general
// (non-diagonal L^+L) collapse operator are Tr(L^+L rho) (hamiltonian.py:97-124 builds the operators).
__global__ void reduced_density_kernel(const c2* psi, long long D, int dim, long long st, double* acc) {
    const long long groups = D / dim;
    double re[16], im[16];
    for (int i = 0; i < 16; ++i) { re[i] = 0.0; im[i] = 0.0; }
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < groups;
         gidx += (long long)gridDim.x * blockDim.x) {
        const long long low = gidx % st, high = gidx / st;
        const long long idx0 = low + high * st * dim;
        c2 v[4];
        for (int a = 0; a < dim; ++a) v[a] = psi[idx0 + a * st];
        for (int a = 0; a < dim; ++a)
            for (int b = 0; b < dim; ++b) {
                re[a * dim + b] = fma(v[a].x, v[b].x, fma(v[a].y, v[b].y, re[a * dim + b]));
                im[a * dim + b] = fma(v[a].y, v[b].x, fma(-v[a].x, v[b].y, im[a * dim + b]));
            }
    }
    for (int i = 0; i < dim * dim; ++i) {
        double x = re[i], y = im[i];
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, o);
            y += __shfl_xor_sync(0xffffffffu, y, o);
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(acc + 2 * i, x); atomicAdd(acc + 2 * i + 1, y); }
    }
}

__global__ void scale_kernel(c2* psi, long long D, double scale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        c2 v = psi[i];
        psi[i] = {v.x * scale, v.y * scale};
    }
}

// y = alpha*y + beta*x (Richardson combination of the step-doubling pair)
__global__ void axpby_kernel(c2* y, const c2* x, double alpha, double beta, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const c2 a = y[i], b = x[i];
        y[i] = {alpha * a.x + beta * b.x, alpha * a.y + beta * b.y};
    }
}

// squared distance per trajectory (step-doubling error estimate)
__global__ void diffnorm2_kernel(const c2* a, const c2* b, long long D, double* out) {
    const long long traj = blockIdx.y;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D;
         i += (long long)gridDim.x * blockDim.x) {
        const c2 x = a[traj * D + i], y = b[traj * D + i];
        const double dr = x.x - y.x, di = x.y - y.y;
        acc = fma(dr, dr, acc);
        acc = fma(di, di, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += ws[i];
        atomicAdd(out + traj, s);
    }
}

}  // namespace pb200

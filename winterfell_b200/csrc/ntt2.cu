// ntt2.cu — the sm_100a NTT pass kernel for sub-transforms of 2^6 .. 2^11 points (ntt.cuh describes the layout,
// the four-step schedule and the parameters; ntt.cu keeps the small-size kernel for sub-transforms below 2^6).
//
// What changed against the round-1 kernel (profiles/r1_ntt_pass_v2_summary.txt: 270-335 issued instructions per
// element per pass, 35 % of them not arithmetic, 20 M shared-memory bank conflicts per LDE launch):
//  * every block owns a tile of 8192 elements (64 KB): S rows x LANES words with LANES = 8 for S <= 2^10 and
//    LANES = 4 (half a segment row) for S = 2^11, so two blocks are resident per SM at every size and the
//    global load / store phases of one overlap the butterflies of the other;
//  * sixteen values per thread per round: a radix-16 butterfly on one lane (64-bit shared-memory accesses) or a
//    radix-8 butterfly on two adjacent lanes (128-bit accesses, one twiddle load for both columns). Round radices
//    per size: 2^11 = 8.16.16, 2^10 = 8.8.16, 2^9 = 8.8.8, 2^8 = 16.16, 2^7 = 8.16, 2^6 = 8.8 — two shared-memory
//    exchanges and two rounds of 64x64-bit twiddle products for 2^10 / 2^11 instead of three / three;
//  * the sub-transform size is a template parameter: all index arithmetic of the rounds is resolved at compile time;
//  * global accesses are 128-bit (two adjacent words of a segment row);
//  * rows are placed in shared memory at (row ^ hash(row)), hash = XOR-fold of the row index above the bits that
//    select the 128-byte bank window, so that rows 2^k apart — what every butterfly stage and the bit-reversed
//    write-back touch together — never share a bank;
//  * inter-round twiddles come from per-round tables [lo][position] (positions of one butterfly contiguous, read as
//    128-bit pairs, conflict-free across the warp) staged by TMA bulk copies; the strided four-step twiddle is
//    w^(e_hi * 2^h) * w^(e_lo) from two small tables instead of a gather over a table of N/2 entries.
#include "minidft.cuh"
#include "ntt.cuh"

#ifndef NTT2_THREADS
#define NTT2_THREADS 256
#endif
#ifndef NTT2_MINB
#define NTT2_MINB 2
#endif
#define NTT2_TILE_ELEMS 8192

// ---- compile-time plan of a sub-transform of 2^LOGS points ----------------------------------------
template <int LOGS>
struct Plan {
    static constexpr int rounds = LOGS >= 9 ? 3 : 2;
    // radix-8 rounds first: the last round carries no twiddles, so the widest butterflies go where they save most
    static constexpr int r0 = (LOGS == 8) ? 4 : 3;
    static constexpr int r1 = (LOGS == 11 || LOGS == 8 || LOGS == 7) ? 4 : 3;
    static constexpr int r2 = LOGS >= 9 ? LOGS - r0 - r1 : 0;
    static_assert(LOGS >= 6 && LOGS <= 11, "plan covers 2^6 .. 2^11");
    static_assert(r0 + r1 + r2 == LOGS && (rounds == 2 || (r2 == 3 || r2 == 4)), "radices must add up");
    static constexpr int lanes = LOGS == 11 ? 4 : 8;
    // twiddle table of round k (k < rounds - 1): [span_k][2^r_k] entries
    static constexpr int tw0_entries = 1 << LOGS;                  // span0 * 2^r0 = S
    static constexpr int tw1_entries = rounds == 3 ? (1 << (LOGS - r0)) : 0;
    static constexpr int tw_entries = tw0_entries + tw1_entries;
};
template <int LOGS> __host__ __device__ constexpr int plan_radix(int k) { return k == 0 ? Plan<LOGS>::r0 : (k == 1 ? Plan<LOGS>::r1 : Plan<LOGS>::r2); }
template <int LOGS> __host__ __device__ constexpr int plan_stage(int k) { return k == 0 ? 0 : (k == 1 ? Plan<LOGS>::r0 : Plan<LOGS>::r0 + Plan<LOGS>::r1); }

// ---- shared-memory row placement --------------------------------------------------------------------
// LK = log2(rows per 128-byte bank window) = 1 for 64-byte rows (LANES = 8), 2 for 32-byte rows (LANES = 4).
template <int LK>
__host__ __device__ constexpr u32 swz_const(u32 r) {
    u32 x = r >> LK, h = 0;
    while (x) { h ^= x & ((1u << LK) - 1); x >>= LK; }
    return h;
}
template <int LK>
__device__ __forceinline__ u32 swz_hash(u32 r) {  // r < 2^11
    if (LK == 1) return __popc(r >> 1) & 1;
    u32 x = r >> 2;           // 9 bits
    x ^= x >> 4;              // bits 0-3 fold 4-7; bit 8 folds into bit 4 -> handled below
    x ^= x >> 8;
    x ^= x >> 2;
    return x & 3;
}
template <int LK>
__device__ __forceinline__ u32 prow(u32 r) { return r ^ swz_hash<LK>(r); }

// ---- TMA bulk copy (global -> shared) -----------------------------------------------------------------
__device__ __forceinline__ void bulk_init(u64* mbar, int tid) {
    if (tid == 0) {
        u32 mb = (u32)__cvta_generic_to_shared(mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
}
__device__ __forceinline__ void bulk_expect(u64* mbar, u32 bytes) {
    u32 mb = (u32)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* smem_dst, const void* gsrc, u32 bytes, u64* mbar) {
    u32 mb = (u32)__cvta_generic_to_shared(mbar), dst = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gsrc), "r"(bytes),
                 "r"(mb)
                 : "memory");
}
__device__ __forceinline__ void bulk_wait(u64* mbar) {
    u32 mb = (u32)__cvta_generic_to_shared(mbar), done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb) : "memory");
}

// ---- one round ----------------------------------------------------------------------------------------
// Round K of the plan on the tile `s` (S rows x LANES words, swizzled rows). tw: this round's table
// [span][2^R] (null for the last round). Each task = 16 values.
template <int LOGS, int K>
__device__ __forceinline__ void tile_round(u64* __restrict__ s, const u64* __restrict__ tw, int tid) {
    constexpr int LANES = Plan<LOGS>::lanes, LK = LANES == 8 ? 1 : 2;
    constexpr int R = plan_radix<LOGS>(K), ST = plan_stage<LOGS>(K);
    constexpr int LOGSPAN = LOGS - ST - R;
    constexpr u32 SPAN = 1u << LOGSPAN;
    constexpr bool LAST = LOGSPAN == 0;
    constexpr u32 TASKS = (1u << LOGS) * LANES / 16;
    if (R == 4) {
        // radix-16 on one lane
        for (u32 task = tid; task < TASKS; task += NTT2_THREADS) {
            const u32 lane = task % LANES, bf = task / LANES;
            const u32 lo = bf & (SPAN - 1), base = ((bf >> LOGSPAN) << (LOGSPAN + 4)) + lo;
            const u32 hb = swz_hash<LK>(base);
            u64 x[16];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const u32 row = (base + ((u32)q << LOGSPAN)) ^ (hb ^ swz_const<LK>((u32)q << LOGSPAN));
                x[q] = s[row * LANES + lane];
            }
            mini_dft<4>(x);
            if (!LAST) {
                const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(tw + (size_t)lo * 16);
#pragma unroll
                for (int h = 0; h < 8; h++) {
                    const ulonglong2 w = t2[h];
                    if (h > 0) x[2 * h] = gl_mul(x[2 * h], w.x);
                    x[2 * h + 1] = gl_mul(x[2 * h + 1], w.y);
                }
            }
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const u32 row = (base + ((u32)q << LOGSPAN)) ^ (hb ^ swz_const<LK>((u32)q << LOGSPAN));
                s[row * LANES + lane] = x[q];
            }
        }
    } else {
        // radix-8 on two adjacent lanes
        constexpr int LP = LANES / 2;
        for (u32 task = tid; task < TASKS; task += NTT2_THREADS) {
            const u32 lp = task % LP, bf = task / LP;
            const u32 lo = bf & (SPAN - 1), base = ((bf >> LOGSPAN) << (LOGSPAN + 3)) + lo;
            const u32 hb = swz_hash<LK>(base);
            u64 xa[8], xb[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const u32 row = (base + ((u32)q << LOGSPAN)) ^ (hb ^ swz_const<LK>((u32)q << LOGSPAN));
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(s + row * LANES + 2 * lp);
                xa[q] = v.x;
                xb[q] = v.y;
            }
            mini_dft<3>(xa);
            mini_dft<3>(xb);
            if (!LAST) {
                const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(tw + (size_t)lo * 8);
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const ulonglong2 w = t2[h];
                    if (h > 0) { xa[2 * h] = gl_mul(xa[2 * h], w.x); xb[2 * h] = gl_mul(xb[2 * h], w.x); }
                    xa[2 * h + 1] = gl_mul(xa[2 * h + 1], w.y);
                    xb[2 * h + 1] = gl_mul(xb[2 * h + 1], w.y);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const u32 row = (base + ((u32)q << LOGSPAN)) ^ (hb ^ swz_const<LK>((u32)q << LOGSPAN));
                *reinterpret_cast<ulonglong2*>(s + row * LANES + 2 * lp) = make_ulonglong2(xa[q], xb[q]);
            }
        }
    }
}

// ---- the pass kernel ------------------------------------------------------------------------------------
// grid = (tile columns x chunks per row, segments, batch). Shared memory:
//   tile [S][LANES] | round twiddles [tw_entries] | post twiddles [S][T] (STRIDED with has_post) | mbarrier
template <int MODE, int LOGS>
__global__ void __launch_bounds__(NTT2_THREADS, NTT2_MINB) ntt2_pass_kernel(const NttPassParams p) {
    extern __shared__ __align__(16) u64 smem[];
    constexpr int LANES = Plan<LOGS>::lanes, LK = LANES == 8 ? 1 : 2, LP = LANES / 2;
    constexpr u32 S = 1u << LOGS;
    const int tid = threadIdx.x;
    const int W = p.W, logW = 31 - __clz(W);
    const int T = W >= LANES ? 1 : (LANES >> logW);         // tile columns per tile
    const int chunks = W >= LANES ? W / LANES : 1;           // tiles across one segment row
    u64* s = smem;
    u64* rtw = s + (size_t)S * LANES;
    u64* ctw = rtw + Plan<LOGS>::tw_entries;
    u64* mbar = ctw + (p.has_post ? (size_t)S * T : 0);

    const u32 tile = blockIdx.x / chunks, q0 = (blockIdx.x % chunks) * LANES;
    const u32 g = blockIdx.y, b = blockIdx.z;
    const u32 R = 1u << p.logR, C = 1u << p.logC;
    const u32 ncols = MODE == NTT_STRIDED ? C : R;

    // round twiddles: one TMA bulk copy
    bulk_init(mbar, tid);
    __syncthreads();
    if (tid == 0) {
        bulk_expect(mbar, Plan<LOGS>::tw_entries * 8);
        bulk_copy(rtw, p.sub_tw, Plan<LOGS>::tw_entries * 8, mbar);
    }

    // post twiddles of this tile: ctw[j][t] = w_M^(+-(j a_mul + (batch0 + b) b_mul) c) * ctab[c] * cconst, c = tile*T + t,
    // with w_M^e = hi_tab[e >> split] * lo_tab[e & (2^split - 1)]
    if (p.has_post && T == 1) {
        // one tile column c: ctw[j] = K g^j with g = w_M^(+-a_mul c), K = w_M^(+-(batch0 + b) b_mul c) ctab[c] cconst — a geometric
        // progression in j: every thread starts at j = tid (one table product) and steps by g^NTT2_THREADS, one multiplication
        // per entry instead of the two or three of the direct form below
        const u32 M = 1u << p.logM, lo_mask = (1u << p.tw_split) - 1, c = tile;
        auto wpow = [&](u64 e64) {
            u32 e = (u32)(e64 & (M - 1));
            if (p.inverse && e) e = M - e;
            return gl_mul(p.tw_hi[e >> p.tw_split], p.tw_lo[e & lo_mask]);
        };
        u64 K = wpow((u64)(p.batch0 + b) * p.b_mul * c);
        if (p.ctab) K = gl_mul(K, p.ctab[c < ncols ? c : 0]);
        if (p.cconst != 1) K = gl_mul(K, p.cconst);
        const u64 step = wpow((u64)NTT2_THREADS * p.a_mul * c);
        u64 x = gl_mul(K, wpow((u64)tid * p.a_mul * c));
        for (u32 j = tid; j < S; j += NTT2_THREADS) {
            ctw[j] = x;
            x = gl_mul(x, step);
        }
    } else if (p.has_post) {
        const u32 M = 1u << p.logM, lo_mask = (1u << p.tw_split) - 1;
        for (u32 idx = tid; idx < S * (u32)T; idx += NTT2_THREADS) {
            const u32 j = idx / T, tt = idx % T, c = tile * T + tt;
            u64 e64 = ((u64)j * p.a_mul + (u64)(p.batch0 + b) * p.b_mul) * c;
            u32 e = (u32)(e64 & (M - 1));
            if (p.inverse && e) e = M - e;
            u64 x = gl_mul(p.tw_hi[e >> p.tw_split], p.tw_lo[e & lo_mask]);
            if (p.ctab) x = gl_mul(x, p.ctab[c < ncols ? c : 0]);
            if (p.cconst != 1) x = gl_mul(x, p.cconst);
            ctw[idx] = x;
        }
    }

    // ---- load: thread -> (row i, lane pair); NTT_LD_BATCH independent 128-bit loads in flight ----
    const u64* in = p.in + (size_t)g * p.in_seg_stride + (size_t)b * p.in_batch_stride;
    const u64* pre = p.pre_tab ? p.pre_tab + (size_t)b * p.pre_batch_stride : nullptr;
    {
        const u32 lp = tid % LP;
        const u32 l0 = 2 * lp;                                     // first lane of the pair
        const u32 t0 = W >= LANES ? 0 : (l0 >> logW), t1 = W >= LANES ? 0 : ((l0 + 1) >> logW);
        const u32 w0 = W >= LANES ? q0 + l0 : (l0 & (W - 1)), w1 = W >= LANES ? q0 + l0 + 1 : ((l0 + 1) & (W - 1));
        const u32 c0 = tile * T + t0, c1 = tile * T + t1;
        const bool ok0 = c0 < ncols, ok1 = c1 < ncols;
        // element (i, col c, word w): STRIDED row C*i + c; CONTIG row c*C + i
        const u64 istride = MODE == NTT_STRIDED ? ((u64)W << p.logC) : (u64)W;
        const u64* a0 = MODE == NTT_STRIDED ? in + (size_t)c0 * W + w0 : in + (((size_t)c0 << p.logC) * W + w0);
        const u64* a1 = MODE == NTT_STRIDED ? in + (size_t)c1 * W + w1 : in + (((size_t)c1 << p.logC) * W + w1);
        const bool vec = p.vec_in;
        constexpr u32 ROWS_PER_IT = NTT2_THREADS / LP;
        constexpr int LDB = 8;
#pragma unroll 1
        for (u32 i0 = tid / LP; i0 < S; i0 += ROWS_PER_IT * LDB) {
            u64 va[LDB], vb[LDB], f[LDB];
#pragma unroll
            for (int k = 0; k < LDB; k++) {
                const u32 i = i0 + k * ROWS_PER_IT;
                va[k] = 0; vb[k] = 0; f[k] = 1;
                if (i < S) {
                    if (vec) {
                        if (ok0) { const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(a0 + (u64)i * istride); va[k] = v.x; vb[k] = v.y; }
                    } else {
                        if (ok0) va[k] = a0[(u64)i * istride];
                        if (ok1) vb[k] = a1[(u64)i * istride];
                    }
                    if (pre) f[k] = pre[i];
                }
            }
#pragma unroll
            for (int k = 0; k < LDB; k++) {
                const u32 i = i0 + k * ROWS_PER_IT;
                if (i < S) {
                    if (pre) { va[k] = gl_mul(va[k], f[k]); vb[k] = gl_mul(vb[k], f[k]); }
                    *reinterpret_cast<ulonglong2*>(s + prow<LK>(i) * LANES + l0) = make_ulonglong2(va[k], vb[k]);
                }
            }
        }
    }
    bulk_wait(mbar);
    __syncthreads();

    // ---- the sub-transform: forward DIF network, output position pos holds X[bitrev(pos)] ----
    tile_round<LOGS, 0>(s, rtw, tid);
    __syncthreads();
    if (Plan<LOGS>::rounds == 3) {
        tile_round<LOGS, 1>(s, rtw + Plan<LOGS>::tw0_entries, tid);
        __syncthreads();
        tile_round<LOGS, 2>(s, nullptr, tid);
    } else {
        tile_round<LOGS, 1>(s, nullptr, tid);
    }
    __syncthreads();

    // ---- write back ----
    u64* out = p.out + (size_t)g * p.out_seg_stride + (size_t)b * p.out_batch_stride;
    {
        const u32 lp = tid % LP, l0 = 2 * lp;
        const u32 t0 = W >= LANES ? 0 : (l0 >> logW), t1 = W >= LANES ? 0 : ((l0 + 1) >> logW);
        const u32 w0 = W >= LANES ? q0 + l0 : (l0 & (W - 1)), w1 = W >= LANES ? q0 + l0 + 1 : ((l0 + 1) & (W - 1));
        const u32 c0 = tile * T + t0, c1 = tile * T + t1;
        const bool ok0 = c0 < ncols, ok1 = c1 < ncols;
        u64 *o0, *o1, jstride;
        if (MODE == NTT_STRIDED) {  // Y[j][m2] = row j*C + col
            o0 = out + (size_t)c0 * W + w0;
            o1 = out + (size_t)c1 * W + w1;
            jstride = (u64)W << p.logC;
        } else {                     // X[j1 + R*j] -> out row (col + R*j)*mul + b*add
            o0 = out + ((size_t)c0 * p.out_row_mul + (size_t)b * p.out_row_add) * p.out_W + p.out_col0 + w0;
            o1 = out + ((size_t)c1 * p.out_row_mul + (size_t)b * p.out_row_add) * p.out_W + p.out_col0 + w1;
            jstride = ((u64)p.out_row_mul << p.logR) * p.out_W;
        }
        const bool post = p.has_post, scale = !p.has_post && p.cconst != 1, vec = p.vec_out;
        const u32 inv_mask = p.inverse ? (S - 1) : 0;  // jf = inverse ? (S - j) mod S : j
        constexpr u32 ROWS_PER_IT = NTT2_THREADS / LP;
        const bool scatter = MODE == NTT_CONTIG && p.sc_on;
        const u32 sc_k = p.sc_coset + b, sc_mask = (1u << p.sc_log_nj) - 1;
        const size_t sc_off = (size_t)(p.sc_seg0 + g) * p.sc_seg_stride + p.out_col0 + w0;
#pragma unroll 4
        for (u32 j = tid / LP; j < S; j += ROWS_PER_IT) {
            const u32 jf = inv_mask ? ((S - j) & inv_mask) : j;
            const u32 r = __brev(jf) >> (32 - LOGS);
            ulonglong2 v = *reinterpret_cast<const ulonglong2*>(s + prow<LK>(r) * LANES + l0);
            if (scatter) {  // no post factor on this path (CONTIG passes of the LDE carry none)
                const u32 jj = c0 + (j << p.logR);
                u64* dst = p.sc_peer[jj >> p.sc_log_nj] + sc_off + (((size_t)(jj & sc_mask) << p.sc_log_b) + sc_k) * p.out_W;
                if (ok0) *reinterpret_cast<ulonglong2*>(dst) = v;
                if (p.sc_world && (jj & sc_mask) == 0 && ok0) {   // halo copy for the rank before the owner
                    const u32 q = jj >> p.sc_log_nj, qp = q ? q - 1 : p.sc_world - 1;
                    *reinterpret_cast<ulonglong2*>(p.sc_peer[qp] + sc_off + ((((size_t)sc_mask + 1) << p.sc_log_b) + sc_k) * p.out_W) = v;
                }
                continue;
            }
            if (post) {
                const u64 f0 = ctw[j * T + t0];
                v.x = gl_mul(v.x, f0);
                v.y = gl_mul(v.y, t1 == t0 ? f0 : ctw[j * T + t1]);
            } else if (scale) {
                v.x = gl_mul(v.x, p.cconst);
                v.y = gl_mul(v.y, p.cconst);
            }
            if (vec) {
                if (ok0) *reinterpret_cast<ulonglong2*>(o0 + (u64)j * jstride) = v;
            } else {
                if (ok0) o0[(u64)j * jstride] = v.x;
                if (ok1) o1[(u64)j * jstride] = v.y;
            }
        }
    }
}

// ---- round-twiddle tables -----------------------------------------------------------------------------
// table of a 2^LOGS-point plan: round k < rounds-1: [lo < span_k][qo < 2^r_k] = w_S^((lo * bitrev_r(qo)) << stage_k)
template <int LOGS>
__global__ void ntt2_build_tw_kernel(u64* out, u64 w_s) {
    const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (u32)Plan<LOGS>::tw_entries) return;
    int k = idx < (u32)Plan<LOGS>::tw0_entries ? 0 : 1;
    const u32 li = k == 0 ? idx : idx - Plan<LOGS>::tw0_entries;
    const int r = plan_radix<LOGS>(k), st = plan_stage<LOGS>(k);
    const u32 lo = li >> r, qo = li & ((1u << r) - 1);
    const u32 e = (lo * (__brev(qo) >> (32 - r))) << st;
    out[idx] = gl_pow(w_s, e);
}

template <int LOGS>
static size_t smem_bytes_t(const NttPassParams& p) {
    const int lanes = Plan<LOGS>::lanes;
    const size_t S = (size_t)1 << LOGS, T = p.W >= lanes ? 1 : lanes / p.W;
    return (S * lanes + Plan<LOGS>::tw_entries + (p.has_post ? S * T : 0) + 2) * 8;
}
template <int MODE, int LOGS>
static cudaError_t launch_t(const NttPassParams& p, u32 n_segments, u32 n_batch, cudaStream_t st) {
    const int lanes = Plan<LOGS>::lanes;
    const u32 T = p.W >= lanes ? 1 : lanes / p.W, chunks = p.W >= lanes ? p.W / lanes : 1;
    const u32 ncols = MODE == NTT_STRIDED ? (1u << p.logC) : (1u << p.logR);
    const size_t smem = smem_bytes_t<LOGS>(p);
    dim3 grid(((ncols + T - 1) / T) * chunks, n_segments, n_batch);
    cudaError_t e = cudaFuncSetAttribute(ntt2_pass_kernel<MODE, LOGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ntt2_pass_kernel<MODE, LOGS><<<grid, NTT2_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}
template <int MODE>
static cudaError_t launch_m(const NttPassParams& p, u32 n_segments, u32 n_batch, cudaStream_t st) {
    switch (p.logS) {
        case 6: return launch_t<MODE, 6>(p, n_segments, n_batch, st);
        case 7: return launch_t<MODE, 7>(p, n_segments, n_batch, st);
        case 8: return launch_t<MODE, 8>(p, n_segments, n_batch, st);
        case 9: return launch_t<MODE, 9>(p, n_segments, n_batch, st);
        case 10: return launch_t<MODE, 10>(p, n_segments, n_batch, st);
        case 11: return launch_t<MODE, 11>(p, n_segments, n_batch, st);
    }
    return cudaErrorInvalidValue;
}
cudaError_t ntt2_launch_pass(int mode, const NttPassParams& p, u32 n_segments, u32 n_batch, cudaStream_t st) {
    return mode == NTT_STRIDED ? launch_m<NTT_STRIDED>(p, n_segments, n_batch, st) : launch_m<NTT_CONTIG>(p, n_segments, n_batch, st);
}
size_t ntt2_tw_entries(int logS) {
    switch (logS) {
        case 6: return Plan<6>::tw_entries; case 7: return Plan<7>::tw_entries; case 8: return Plan<8>::tw_entries;
        case 9: return Plan<9>::tw_entries; case 10: return Plan<10>::tw_entries; case 11: return Plan<11>::tw_entries;
    }
    return 0;
}
cudaError_t ntt2_build_tw(int logS, u64* d_out, cudaStream_t st) {
    const u64 w = gl_root_of_unity((u32)logS);
    const unsigned nb = (unsigned)((ntt2_tw_entries(logS) + 255) / 256);
    switch (logS) {
        case 6: ntt2_build_tw_kernel<6><<<nb, 256, 0, st>>>(d_out, w); break;
        case 7: ntt2_build_tw_kernel<7><<<nb, 256, 0, st>>>(d_out, w); break;
        case 8: ntt2_build_tw_kernel<8><<<nb, 256, 0, st>>>(d_out, w); break;
        case 9: ntt2_build_tw_kernel<9><<<nb, 256, 0, st>>>(d_out, w); break;
        case 10: ntt2_build_tw_kernel<10><<<nb, 256, 0, st>>>(d_out, w); break;
        case 11: ntt2_build_tw_kernel<11><<<nb, 256, 0, st>>>(d_out, w); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

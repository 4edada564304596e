// ntt.cu — see ntt.cuh for the layout and schedule.
#include "ntt.cuh"

#include "minidft.cuh"

// One radix-2^R round at DIF stage `stage` on the S x 8 tile `s` (row-major, 8 lanes per row).
template <int R>
__device__ __forceinline__ void dif_round(u64* s, const u64* stw, int logS, int stage, int tid) {
    const int logspan = logS - stage - R;
    const u32 span = 1u << logspan;
    const u32 nbf = (1u << logS) >> R;
    const int lane = tid & (NTT_LANES - 1);
    for (u32 bf = tid >> 3; bf < nbf; bf += NTT_THREADS / NTT_LANES) {
        u32 lo = bf & (span - 1);
        u32 base = ((bf >> logspan) << (logspan + R)) + lo;
        u64 x[1 << R];
#pragma unroll
        for (int q = 0; q < (1 << R); q++) x[q] = s[((base + ((u32)q << logspan)) << 3) + lane];
        mini_dft<R>(x);
        if (logspan > 0) {
#pragma unroll
            for (int qo = 1; qo < (1 << R); qo++) {
                // position qo holds frequency bitrev(qo) of the mini-DFT: twiddle w_G^(lo * freq)
                u32 e = (lo * brev(qo, R)) << stage;   // < S: the table holds w_S^e for all e in [0, S)
                x[qo] = gl_mul(x[qo], stw[e]);
            }
        }
#pragma unroll
        for (int q = 0; q < (1 << R); q++) s[((base + ((u32)q << logspan)) << 3) + lane] = x[q];
    }
}

__device__ __forceinline__ void tile_dft(u64* s, const u64* stw, int logS, int tid) {
    int stage = 0;
#ifdef NTT_RADIX32
    while (logS - stage >= 5) {
        dif_round<5>(s, stw, logS, stage, tid);
        stage += 5;
        __syncthreads();
    }
#endif
#ifdef NTT_RADIX16
    while (logS - stage >= 4) {
        dif_round<4>(s, stw, logS, stage, tid);
        stage += 4;
        __syncthreads();
    }
#endif
    while (logS - stage >= 3) {
        dif_round<3>(s, stw, logS, stage, tid);
        stage += 3;
        __syncthreads();
    }
    if (logS - stage == 2) {
        dif_round<2>(s, stw, logS, stage, tid);
        __syncthreads();
    } else if (logS - stage == 1) {
        dif_round<1>(s, stw, logS, stage, tid);
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------------------------
// TMA bulk copy of the sub-transform twiddle table into shared memory.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gsrc, u32 bytes, u64* mbar, int tid) {
    u32 mb = (u32)__cvta_generic_to_shared(mbar);
    u32 dst = (u32)__cvta_generic_to_shared(smem_dst);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
            "l"(gsrc), "r"(bytes), "r"(mb)
            : "memory");
    }
}
__device__ __forceinline__ void tma_bulk_wait(u64* mbar) {
    u32 mb = (u32)__cvta_generic_to_shared(mbar);
    u32 done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mb)
            : "memory");
    }
}

// -------------------------------------------------------------------------------------------------
// The pass kernel. grid = (tiles, segments, batch), block = NTT_THREADS.
// Shared memory: tile [S][8] | sub twiddles [S/2] | post twiddles [S][T] (if has_post) | mbarrier
// -------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) ntt_pass_kernel(const NttPassParams p) {
    extern __shared__ __align__(16) u64 smem[];
    const int tid = threadIdx.x;
    const int logS = p.logS;
    const u32 S = 1u << logS;
    const int W = p.W;
    const int logW = 31 - __clz(W);
    const int T = NTT_LANES >> logW;
    u64* s = smem;
    u64* stw = s + (size_t)S * NTT_LANES;   // S entries: w_S^e, e < S (second half = negated first half)
    u64* ctw = stw + S;
    u64* mbar = ctw + (p.has_post ? (size_t)S * T : 0);

    const u32 tile = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
    const u32 R = 1u << p.logR, C = 1u << p.logC;
    const int lane = tid & 7;
    const int t = lane >> logW;        // tile column of this lane
    const int q = lane & (W - 1);      // segment column of this lane
    const u32 col = tile * T + t;      // STRIDED: m2; CONTIG: j1
    const bool col_ok = (MODE == NTT_STRIDED) ? (col < C) : (col < R);

    // stage the twiddle table with one bulk copy (needs >= 16 bytes, 16-byte aligned)
    const u32 tw_bytes = (S >> 1) * 8;
    const bool use_tma = tw_bytes >= 16;
    if (use_tma) tma_bulk_load(stw, p.sub_tw, tw_bytes, mbar, tid);
    else if (tid < (int)(S >> 1)) stw[tid] = p.sub_tw[tid];

    // post twiddles for this tile: ctw[j][t] (gathers batched the same way)
    if (p.has_post) {
        const u32 M = 1u << p.logM, Mh = M >> 1;
        const u32 total = S * (u32)T;
        const int logT = 3 - logW;
        for (u32 idx0 = tid; idx0 < total; idx0 += NTT_THREADS * 4) {
            u64 w[4], cc[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 idx = idx0 + k * NTT_THREADS;
                w[k] = 1;
                cc[k] = 1;
                if (idx < total) {
                    u32 j = idx >> logT, tt = idx & (T - 1);
                    u32 c = tile * T + tt;
                    u64 e64 = ((u64)j * p.a_mul + (u64)(p.batch0 + b) * p.b_mul) * c;
                    u32 e = (u32)(e64 & (M - 1));
                    if (p.inverse && e) e = M - e;
                    w[k] = tw_lookup(p.master, e, Mh);
                    if (p.ctab) cc[k] = p.ctab[c < (MODE == NTT_STRIDED ? C : R) ? c : 0];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 idx = idx0 + k * NTT_THREADS;
                if (idx < total) {
                    u64 x = w[k];
                    if (p.ctab) x = gl_mul(x, cc[k]);
                    if (p.cconst != 1) x = gl_mul(x, p.cconst);
                    ctw[idx] = x;
                }
            }
        }
    }

    // load the tile: batches of NTT_LD_BATCH independent loads per thread are issued before any is
    // consumed (the first profile showed long-scoreboard stalls of 2.7 cycles per issued instruction
    // from one-at-a-time LDG -> STS chains)
    const u64* in = p.in + (size_t)g * p.in_seg_stride + (size_t)b * p.in_batch_stride;
    const u64* pre = p.pre_tab ? p.pre_tab + (size_t)b * p.pre_batch_stride : nullptr;
    constexpr u32 ROWS_PER_IT = NTT_THREADS / NTT_LANES;
    if (MODE == NTT_CONTIG && W < NTT_LANES) {
        // narrow segments: lane = (tile column t, segment column q) and every tile column is its own
        // contiguous run of C*W words, so map threads along the run (index-fastest) instead of along the
        // lanes: a warp then reads 256 contiguous bytes instead of 8-byte pieces of 8 different rows
        const u32 run = S << logW;  // words per tile column
        for (u32 e0 = tid; e0 < run * (u32)T; e0 += NTT_THREADS * NTT_LD_BATCH) {
            u64 v[NTT_LD_BATCH], f[NTT_LD_BATCH];
#pragma unroll
            for (int k = 0; k < NTT_LD_BATCH; k++) {
                u32 e = e0 + k * NTT_THREADS;
                v[k] = 0;
                f[k] = 1;
                if (e < run * (u32)T) {
                    u32 tt = e / run, r = e % run, i = r >> logW;
                    u32 c = tile * T + tt;
                    if (c < R) {
                        v[k] = in[(((size_t)c << p.logC) << logW) + r];
                        if (pre) f[k] = pre[i];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NTT_LD_BATCH; k++) {
                u32 e = e0 + k * NTT_THREADS;
                if (e < run * (u32)T) {
                    u32 tt = e / run, r = e % run, i = r >> logW, qq = r & (W - 1);
                    s[(i << 3) + (tt << logW) + qq] = pre ? gl_mul(v[k], f[k]) : v[k];
                }
            }
        }
    } else {
        // the address is linear in the tile row i: one 64-bit base per thread, one multiply-add per load
        // (STRIDED: row C*m1 + m2, i = m1; CONTIG: row j1*C + m2, i = m2)
        const u64* in_thr = MODE == NTT_STRIDED ? in + (size_t)col * W + q : in + (((size_t)col << p.logC) * W + q);
        const u64 istride = MODE == NTT_STRIDED ? ((u64)W << p.logC) : (u64)W;
        u64* s_thr = s + lane;
        for (u32 i0 = tid >> 3; i0 < S; i0 += ROWS_PER_IT * NTT_LD_BATCH) {
            u64 v[NTT_LD_BATCH], f[NTT_LD_BATCH];
#pragma unroll
            for (int k = 0; k < NTT_LD_BATCH; k++) {
                u32 i = i0 + k * ROWS_PER_IT;
                v[k] = 0;
                f[k] = 1;
                if (col_ok && i < S) {
                    v[k] = in_thr[(u64)i * istride];
                    if (pre) f[k] = pre[i];
                }
            }
#pragma unroll
            for (int k = 0; k < NTT_LD_BATCH; k++) {
                u32 i = i0 + k * ROWS_PER_IT;
                if (i < S) s_thr[i << 3] = pre ? gl_mul(v[k], f[k]) : v[k];
            }
        }
    }
    if (use_tma) tma_bulk_wait(mbar);
    __syncthreads();
    for (u32 e = tid; e < (S >> 1); e += NTT_THREADS) stw[(S >> 1) + e] = gl_neg(stw[e]);  // w^(e + S/2) = -w^e
    __syncthreads();

    tile_dft(s, stw, logS, tid);

    // write back: smem position pos holds forward X[bitrev(pos)]
    u64* out = p.out + (size_t)g * p.out_seg_stride + (size_t)b * p.out_batch_stride;
    if (col_ok) {
        // linear in j: STRIDED Y[j][m2] = row j*C + col; CONTIG X[j1 + R*j] -> out row (col + R*j)*mul + b*add
        u64* out_thr;
        u64 jstride;
        if (MODE == NTT_STRIDED) {
            out_thr = out + (size_t)col * W + q;
            jstride = (u64)W << p.logC;
        } else {
            out_thr = out + ((size_t)col * p.out_row_mul + (size_t)b * p.out_row_add) * p.out_W + p.out_col0 + q;
            jstride = ((u64)p.out_row_mul << p.logR) * p.out_W;
        }
        const u64* s_thr = s + lane;
        const u64* ctw_thr = ctw + t;
        const bool post = p.has_post, scale = !p.has_post && p.cconst != 1;
        const u32 inv_mask = p.inverse ? (S - 1) : 0;  // jf = inverse ? (S - j) mod S : j
        for (u32 j = tid >> 3; j < S; j += NTT_THREADS / NTT_LANES) {
            u32 jf = inv_mask ? ((S - j) & inv_mask) : j;
            u64 v = s_thr[brev(jf, logS) << 3];
            if (post) v = gl_mul(v, ctw_thr[j * T]);
            else if (scale) v = gl_mul(v, p.cconst);
            out_thr[(u64)j * jstride] = v;
        }
    }
}

size_t ntt_pass_smem_bytes(const NttPassParams& p) {
    size_t S = (size_t)1 << p.logS;
    size_t T = NTT_LANES / p.W;
    size_t words = S * NTT_LANES + S + (p.has_post ? S * T : 0) + 2;
    return words * 8;
}

cudaError_t ntt_launch_pass(int mode, const NttPassParams& p, u32 n_segments, u32 n_batch, cudaStream_t st) {
    size_t smem = ntt_pass_smem_bytes(p);
    u32 T = NTT_LANES / p.W;
    u32 ncols = mode == NTT_STRIDED ? (1u << p.logC) : (1u << p.logR);
    dim3 grid((ncols + T - 1) / T, n_segments, n_batch);
    cudaError_t e;
    if (mode == NTT_STRIDED) {
        e = cudaFuncSetAttribute(ntt_pass_kernel<NTT_STRIDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ntt_pass_kernel<NTT_STRIDED><<<grid, NTT_THREADS, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(ntt_pass_kernel<NTT_CONTIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        ntt_pass_kernel<NTT_CONTIG><<<grid, NTT_THREADS, smem, st>>>(p);
    }
    return cudaGetLastError();
}

/*
 * oracle/moe_oracle.c -- CPU restatement of the reference's fused MoE forward semantics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by the product
 * (flashmoe_b200/, include/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / reported CPU baseline.
 *
 * PARITY UNPINNED: the reference (osayamenja/FlashMoE @ 1a77aed4) ships no golden vectors, known-answer
 * tests or fixtures for this path and cannot be compiled or run without a GPU plus
 * Boost/MatX/cuBLASDx/NVSHMEM, so this restatement is anchored on the reference
 * SOURCE only.  Each function cites the reference file:line it follows (paths relative to the reference
 * root).  Arithmetic that lives in un-vendored third-party code (CUTLASS `main` CollectiveMma / NumericConverter
 * / GELU, CUB BlockScan) is restated from its published behaviour: fp32 accumulation of exact bf16 x bf16
 * products, round-to-nearest-even float->bf16, erf-form GELU, integer inclusive prefix sums.
 *
 * Numeric model:
 *   - bf16 operands, fp32 accumulation (sequential, ascending k), RNE at the three store points
 *     (gateOut, h, y) -- csrc/include/flashmoe/os/processor/processor.cuh:423, gemm.cuh:24-32.
 *   - softmax in fp32 with the reference's ONLINE recurrence (moe/gate.cuh:575-584) and a model of the
 *     fast intrinsics it names: __expf(z) = ex2.approx.ftz(z * log2e)  ->  exp2f(fl(z*log2e)), flushed to 0
 *     below 2^-126;  __fdividef(a,b) -> a / b (true division; the approx reciprocal differs by <=2 ulp).
 *   - top-k on fp32 p with strict '>' scanning experts in ascending order (moe/gate.cuh:654-670).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define FMO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ bf16 */
static inline float bf16_to_f32(uint16_t v) {
    uint32_t u = ((uint32_t)v) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* round-to-nearest-even, NaN preserved (cutlass::NumericConverter<bfloat16_t,float>) */
static inline uint16_t f32_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}

static inline float rne_bf16(float f) { return bf16_to_f32(f32_to_bf16(f)); }

FMO_API void fmo_bf16_to_f32(const uint16_t* src, float* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = bf16_to_f32(src[i]);
}
FMO_API void fmo_f32_to_bf16(const float* src, uint16_t* dst, int64_t n) {
    for (int64_t i = 0; i < n; ++i) dst[i] = f32_to_bf16(src[i]);
}

FMO_API int fmo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
FMO_API void fmo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* model of __expf under the reference's pip build (--use_fast_math => ftz): moe/gate.cuh:580-583 */
static inline float fast_exp_model(float z) {
    const float t = z * 1.4426950408889634f; /* fl(z * log2 e), one fp32 rounding like the MUFU path */
    if (t < -126.0f) return 0.0f;            /* result would be denormal -> flushed */
    return exp2f(t);
}

/* ------------------------------------------------------------------ GEMM core
 * C[m][n] = sum_k A[m][k] * B[n][k]   (A row-major [M,K], B "K-major" [N,K] like W_up[P,H] / W_down[H,P],
 * moe/moe.cuh:111-116).  fp32 accumulation, ascending k for every output element.  B is transposed once
 * into Bt[K][N] so the inner loop runs over n (vectorises without reassociating the k-sum).
 * epilogue: v = acc + bias[n]; act; RNE to bf16  (gemm.cuh:24-32 FAA functor; processor.cuh:423).
 */
#define MR 4
#define NR 16
typedef float v8f __attribute__((vector_size(32)));
static inline v8f splat8(float a) { return (v8f){a, a, a, a, a, a, a, a}; }

static void transpose_to_f32(const uint16_t* B, float* Bt, int N, int K) {
#pragma omp parallel for schedule(static)
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int n = 0; n < N; ++n) {
            const uint16_t* row = B + (size_t)n * K;
            int kend = k0 + 32 < K ? k0 + 32 : K;
            for (int k = k0; k < kend; ++k) Bt[(size_t)k * N + n] = bf16_to_f32(row[k]);
        }
    }
}

static inline float act_apply(float v, int act) {
    if (act == 0) return v > 0.0f ? v : 0.0f;                       /* cutlass ReLU, types.cuh:151-159 */
    if (act == 1) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); /* cutlass GELU (erf form) */
    return v;                                                       /* identity (GEMM1, types.cuh:460) */
}

/* A given as bf16 [M,K]; out bf16 [M,N].  act: 0 relu, 1 gelu, 2 identity.  bias may be NULL (bf16 [N]). */
static void gemm_bias_act(const uint16_t* A, const float* Bt, const uint16_t* bias, uint16_t* out, int64_t M,
                          int N, int K, int act) {
    const int64_t mblocks = (M + MR - 1) / MR;
#pragma omp parallel
    {
        float* arow = (float*)malloc(sizeof(float) * (size_t)MR * K);
#pragma omp for schedule(dynamic, 4)
        for (int64_t mb = 0; mb < mblocks; ++mb) {
            const int64_t m0 = mb * MR;
            const int mr = (int)((M - m0) < MR ? (M - m0) : MR);
            for (int r = 0; r < MR; ++r) {
                if (r < mr)
                    for (int k = 0; k < K; ++k) arow[(size_t)r * K + k] = bf16_to_f32(A[(size_t)(m0 + r) * K + k]);
                else
                    for (int k = 0; k < K; ++k) arow[(size_t)r * K + k] = 0.0f;
            }
            for (int n0 = 0; n0 < N; n0 += NR) {
                float acc[MR][NR];
                if (n0 + NR <= N) {
                    /* 4 x 16 register tile: every acc lane is its own ascending-k fp32 sum (no reassociation) */
                    v8f c00 = {0}, c01 = {0}, c10 = {0}, c11 = {0}, c20 = {0}, c21 = {0}, c30 = {0}, c31 = {0};
                    const float *a0 = arow, *a1 = arow + K, *a2 = arow + 2 * (size_t)K, *a3 = arow + 3 * (size_t)K;
                    for (int k = 0; k < K; ++k) {
                        const float* b = Bt + (size_t)k * N + n0;
                        v8f b0, b1;
                        memcpy(&b0, b, 32);
                        memcpy(&b1, b + 8, 32);
                        const v8f x0 = splat8(a0[k]), x1 = splat8(a1[k]), x2 = splat8(a2[k]), x3 = splat8(a3[k]);
                        c00 += x0 * b0; c01 += x0 * b1;
                        c10 += x1 * b0; c11 += x1 * b1;
                        c20 += x2 * b0; c21 += x2 * b1;
                        c30 += x3 * b0; c31 += x3 * b1;
                    }
                    memcpy(&acc[0][0], &c00, 32); memcpy(&acc[0][8], &c01, 32);
                    memcpy(&acc[1][0], &c10, 32); memcpy(&acc[1][8], &c11, 32);
                    memcpy(&acc[2][0], &c20, 32); memcpy(&acc[2][8], &c21, 32);
                    memcpy(&acc[3][0], &c30, 32); memcpy(&acc[3][8], &c31, 32);
                } else {
                    for (int r = 0; r < MR; ++r)
                        for (int j = 0; j < NR; ++j) acc[r][j] = 0.0f;
                    const int nr = N - n0;
                    for (int k = 0; k < K; ++k) {
                        const float* b = Bt + (size_t)k * N + n0;
                        for (int r = 0; r < MR; ++r) {
                            const float a = arow[(size_t)r * K + k];
                            for (int j = 0; j < nr; ++j) acc[r][j] += a * b[j];
                        }
                    }
                }
                const int nr = (N - n0) < NR ? (N - n0) : NR;
                for (int r = 0; r < mr; ++r)
                    for (int j = 0; j < nr; ++j) {
                        float v = acc[r][j];
                        if (bias) v += bf16_to_f32(bias[n0 + j]);
                        out[(size_t)(m0 + r) * N + n0 + j] = f32_to_bf16(act_apply(v, act));
                    }
            }
        }
        free(arow);
    }
}

/* ------------------------------------------------------------------ A.2-A.4 gate
 * x [S,H] bf16; wg_eff [E,H] bf16 = the caller's [H,E] tensor REINTERPRETED flat as [E,H]
 * (python_bindings.cu:93-99 + moe/moe.cuh:107-109 -- a reinterpretation, not a transpose).
 * Outputs: logits f32 [S,E]; probs f32 [S,E]; gate_out bf16 [S,E] (first E columns of the reference's [S,PX]);
 *          topk_idx i32 [S,k]; mcw f32 [S]; abs_sum f32 [S] = max_e sum_h |x*w| (error scale for ambiguity flags).
 */
FMO_API int fmo_gate(const uint16_t* x, const uint16_t* wg_eff, int S, int H, int E, int k, float* logits,
                     float* probs, uint16_t* gate_out, int32_t* topk_idx, float* mcw, float* abs_sum) {
    if (k < 1 || k > E) return -1;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < S; ++t) {
        const uint16_t* xr = x + (size_t)t * H;
        float* l = logits + (size_t)t * E;
        float amax = 0.0f;
        for (int e = 0; e < E; ++e) {
            const uint16_t* wr = wg_eff + (size_t)e * H;
            float acc = 0.0f, aacc = 0.0f;
            for (int h = 0; h < H; ++h) { /* fp32 accumulate, ascending h (moe/gate.cuh:526-535) */
                const float prod = bf16_to_f32(xr[h]) * bf16_to_f32(wr[h]);
                acc += prod;
                aacc += fabsf(prod);
            }
            l[e] = acc;
            if (aacc > amax) amax = aacc;
        }
        abs_sum[t] = amax;
        /* online softmax, exactly the recurrence of moe/gate.cuh:575-584 */
        float dI = 0.0f, mI = -INFINITY;
        for (int e = 0; e < E; ++e) {
            const float pM = mI;
            mI = fmaxf(mI, l[e]);
            dI = fmaf(dI, fast_exp_model(pM - mI), fast_exp_model(l[e] - mI));
        }
        float* p = probs + (size_t)t * E;
        for (int e = 0; e < E; ++e) {
            p[e] = fast_exp_model(l[e] - mI) / dI;
            gate_out[(size_t)t * E + e] = f32_to_bf16(p[e]); /* moe/gate.cuh:590-605 */
        }
        /* k rounds of strict-'>' argmax over unselected experts (moe/gate.cuh:654-670) */
        float sum = 0.0f;
        for (int i = 0; i < k; ++i) {
            float sV = -INFINITY;
            int sIdx = 0;
            for (int j = 0; j < E; ++j) {
                int taken = 0;
                for (int q = 0; q < i; ++q) taken |= (topk_idx[(size_t)t * k + q] == j);
                if (p[j] > sV && !taken) {
                    sIdx = j;
                    sV = p[j];
                }
            }
            topk_idx[(size_t)t * k + i] = sIdx;
            sum += sV;
        }
        mcw[t] = sum;
    }
    return 0;
}

/* ------------------------------------------------------------------ A.5 slotting & capacity
 * Canonical order = ascending token index (a legal interleaving of the reference's per-tile BlockScan +
 * atomicAdd(eC) arrival order, moe/gate.cuh:678-718).  slot[t,j] = position of token t among all tokens that
 * selected expert topk_idx[t,j] (selections are counted even when dropped); kept iff slot < EC
 * (moe/gate.cuh:713-717, os/packet.cuh:85,112-113).  counts[e] = all selections of e.
 */
FMO_API int fmo_slots(const int32_t* topk_idx, int S, int E, int k, int EC, int32_t* slot, int32_t* kept,
                      int32_t* counts) {
    for (int e = 0; e < E; ++e) counts[e] = 0;
    for (int t = 0; t < S; ++t)
        for (int j = 0; j < k; ++j) {
            const int e = topk_idx[(size_t)t * k + j];
            if (e < 0 || e >= E) return -1;
            const int s = counts[e]++;
            slot[(size_t)t * k + j] = s;
            kept[(size_t)t * k + j] = s < EC ? 1 : 0;
        }
    return 0;
}

/* ------------------------------------------------------------------ A.6 expert FFN on a packet of rows
 * rows [R,H] bf16 (the dispatched token rows of ONE expert), w_up [P,H] bf16, w_down_eff [H,P] bf16 (the
 * caller's [P,H] block reinterpreted flat, python_bindings.cu:113-119), biases bf16 or NULL.
 * h = rne(act(rows . w_up^T + b_up)) [R,P]   (processor.cuh:685-695 fGET<PreGEMM>)
 * y = rne(h . w_down_eff^T + b_down)   [R,H]   (processor.cuh:711-721 fGET<PostGEMM>, identity activation)
 */
FMO_API int fmo_expert_ffn(const uint16_t* rows, int64_t R, int H, int P, const uint16_t* w_up,
                           const uint16_t* w_down_eff, const uint16_t* b_up, const uint16_t* b_down, int act,
                           uint16_t* h_out, uint16_t* y_out) {
    if (R <= 0) return 0;
    float* bt = (float*)malloc(sizeof(float) * (size_t)H * P);
    if (!bt) return -2;
    transpose_to_f32(w_up, bt, P, H); /* Bt[H][P] */
    gemm_bias_act(rows, bt, b_up, h_out, R, P, H, act);
    transpose_to_f32(w_down_eff, bt, H, P); /* Bt[P][H] */
    gemm_bias_act(h_out, bt, b_down, y_out, R, H, P, 2);
    free(bt);
    return 0;
}

/* ------------------------------------------------------------------ A.7 combine for one token row
 * k > 1 (CombineMode::multithreaded, processor.cuh:110-169):
 *     term_j = rne( p~_j (x) rne( float(y_j[c]) / mCw ) ),  (x) = bf16 x bf16 multiply (fp32 product, RNE)
 *     out[c] = bf16 accumulation of the kept terms, here in ascending j (for k == 2 the order is immaterial:
 *              0 + a is exact and bf16 addition commutes).
 * k == 1 (CombineMode::single, processor.cuh:170-203): out = y, no scaling; dropped -> zeros.
 * y_rows: k pointers to bf16 [H] (NULL when the pair was dropped); p_tilde: bf16 gateOut[t, e_j].
 */
FMO_API void fmo_combine_token(const uint16_t* const* y_rows, const uint16_t* p_tilde, float mcw, int k, int H,
                               uint16_t* out) {
    if (k == 1) {
        if (y_rows[0])
            memcpy(out, y_rows[0], sizeof(uint16_t) * (size_t)H);
        else
            memset(out, 0, sizeof(uint16_t) * (size_t)H);
        return;
    }
    for (int c = 0; c < H; ++c) {
        float acc = 0.0f; /* output zero-initialised each launch (moe/moe.cuh:43-48) */
        for (int j = 0; j < k; ++j) {
            if (!y_rows[j]) continue;
            const float q = rne_bf16(bf16_to_f32(y_rows[j][c]) / mcw);
            const float term = rne_bf16(bf16_to_f32(p_tilde[j]) * q);
            acc = rne_bf16(acc + term); /* atomicAdd on Element (bf16) */
        }
        out[c] = f32_to_bf16(acc);
    }
}

/* ------------------------------------------------------------------ whole layer, one rank's tokens, all experts
 * local (world size 1 view; the multi-rank composition in oracle/moe_oracle.py calls the pieces above per rank,
 * SURVEY.md A.8).  w_up [E,P,H], w_down_eff [E,H,P], biases [E,P]/[E,H] or NULL.
 * Optional outputs (may be NULL): h_all / y_all [S*k rows in (expert, slot) order are not kept] -- only `out`.
 * Returns 0, or <0 on bad arguments / allocation failure.
 */
FMO_API int fmo_forward(const uint16_t* x, const uint16_t* wg_eff, const uint16_t* w_up,
                        const uint16_t* w_down_eff, const uint16_t* b_up, const uint16_t* b_down, int S, int H,
                        int P, int E, int k, int EC, int act, uint16_t* out, int32_t* topk_idx_out,
                        int32_t* slot_out, int32_t* kept_out, int32_t* counts_out, float* mcw_out,
                        uint16_t* gate_out_out, float* logits_out, float* abs_sum_out) {
    const size_t SE = (size_t)S * E, SK = (size_t)S * k;
    float* logits = logits_out ? logits_out : (float*)malloc(sizeof(float) * SE);
    float* probs = (float*)malloc(sizeof(float) * SE);
    uint16_t* gate_out = gate_out_out ? gate_out_out : (uint16_t*)malloc(sizeof(uint16_t) * SE);
    int32_t* topk = topk_idx_out ? topk_idx_out : (int32_t*)malloc(sizeof(int32_t) * SK);
    int32_t* slot = slot_out ? slot_out : (int32_t*)malloc(sizeof(int32_t) * SK);
    int32_t* kept = kept_out ? kept_out : (int32_t*)malloc(sizeof(int32_t) * SK);
    int32_t* counts = counts_out ? counts_out : (int32_t*)malloc(sizeof(int32_t) * (size_t)E);
    float* mcw = mcw_out ? mcw_out : (float*)malloc(sizeof(float) * (size_t)S);
    float* abs_sum = abs_sum_out ? abs_sum_out : (float*)malloc(sizeof(float) * (size_t)S);
    int rc = -2;
    uint16_t *rows = NULL, *hbuf = NULL, *ybuf = NULL;
    int64_t* yoff = NULL;
    if (!logits || !probs || !gate_out || !topk || !slot || !kept || !counts || !mcw || !abs_sum) goto done;
    rc = fmo_gate(x, wg_eff, S, H, E, k, logits, probs, gate_out, topk, mcw, abs_sum);
    if (rc) goto done;
    rc = fmo_slots(topk, S, E, k, EC, slot, kept, counts);
    if (rc) goto done;
    rc = -2;
    /* per-expert packets: rows in slot order (os/packet.cuh:112-166 copies x[tokenIdx] into heap slot) */
    yoff = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E + 1));
    if (!yoff) goto done;
    yoff[0] = 0;
    for (int e = 0; e < E; ++e) yoff[e + 1] = yoff[e] + (counts[e] < EC ? counts[e] : EC);
    {
        const int64_t R = yoff[E];
        rows = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(R ? R : 1) * H);
        hbuf = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(R ? R : 1) * P);
        ybuf = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(R ? R : 1) * H);
        if (!rows || !hbuf || !ybuf) goto done;
        for (int t = 0; t < S; ++t)
            for (int j = 0; j < k; ++j)
                if (kept[(size_t)t * k + j]) {
                    const int e = topk[(size_t)t * k + j];
                    memcpy(rows + (size_t)(yoff[e] + slot[(size_t)t * k + j]) * H, x + (size_t)t * H,
                           sizeof(uint16_t) * (size_t)H);
                }
        for (int e = 0; e < E; ++e) {
            const int64_t r = yoff[e + 1] - yoff[e];
            rc = fmo_expert_ffn(rows + (size_t)yoff[e] * H, r, H, P, w_up + (size_t)e * P * H,
                                w_down_eff + (size_t)e * H * P, b_up ? b_up + (size_t)e * P : NULL,
                                b_down ? b_down + (size_t)e * H : NULL, act, hbuf + (size_t)yoff[e] * P,
                                ybuf + (size_t)yoff[e] * H);
            if (rc) goto done;
        }
    }
#pragma omp parallel for schedule(static)
    for (int t = 0; t < S; ++t) {
        const uint16_t* yr[16];
        uint16_t pt[16];
        for (int j = 0; j < k; ++j) {
            const int e = topk[(size_t)t * k + j];
            yr[j] = kept[(size_t)t * k + j] ? ybuf + (size_t)(yoff[e] + slot[(size_t)t * k + j]) * H : NULL;
            pt[j] = gate_out[(size_t)t * E + e];
        }
        fmo_combine_token(yr, pt, mcw[t], k, H, out + (size_t)t * H);
    }
    rc = 0;
done:
    free(rows);
    free(hbuf);
    free(ybuf);
    free(yoff);
    free(probs);
    if (!logits_out) free(logits);
    if (!gate_out_out) free(gate_out);
    if (!topk_idx_out) free(topk);
    if (!slot_out) free(slot);
    if (!kept_out) free(kept);
    if (!counts_out) free(counts);
    if (!mcw_out) free(mcw);
    if (!abs_sum_out) free(abs_sum);
    return rc;
}

/* ------------------------------------------------------------------ whole layer on a SAMPLE of one rank's tokens
 * Gate (A.2-A.4) and slots (A.5) run over all S tokens -- capacity drops depend on every earlier token -- but the
 * expert FFN (A.6) and the combine (A.7) only for the n_sample tokens listed in `sample` (ascending or not).  The
 * cost of the FFN is per row, so full-size shapes (d_model 4096, ffn 14336) stay at seconds.  out_sample [n_sample,H].
 * Used by bench.py's post-run parity check and the full-size multi-GPU parity runs.
 */
/* Variant used by the parity checks: `topk_w_given` [n_sample,k] (bf16 bits) / `mcw_given` [n_sample], when not NULL,
 * replace the oracle's own combine weights of the sampled tokens (gateOut[t, e_j] and the top-k probability sum).  The
 * bf16 rounding of a gate probability is decided by the last bits of its fp32 logit, i.e. by the summation order of
 * the router GEMM (sequential here, tensor-core tiles on the device and in the reference): a device weight may sit
 * one bf16 ulp (0.4 %) from the oracle's.  Feeding the device's weights separates that from the expert FFN + combine
 * arithmetic, which is then held to the usual tolerance; the weights themselves are held to <= 1 ulp by the caller. */
FMO_API int fmo_forward_sample_w(const uint16_t* x, const uint16_t* wg_eff, const uint16_t* w_up,
                                 const uint16_t* w_down_eff, const uint16_t* b_up, const uint16_t* b_down, int S, int H,
                                 int P, int E, int k, int EC, int act, const int32_t* sample, int n_sample,
                                 const uint16_t* topk_w_given, const float* mcw_given,
                                 uint16_t* out_sample, int32_t* topk_idx_out, int32_t* slot_out, int32_t* kept_out,
                                 int32_t* counts_out, float* mcw_out, uint16_t* gate_out_out, float* logits_out,
                                 float* abs_sum_out) {
    const size_t SE = (size_t)S * E;
    float* probs = (float*)malloc(sizeof(float) * SE);
    int rc = -2;
    uint16_t *rows = NULL, *hbuf = NULL, *ybuf = NULL;
    int64_t *yoff = NULL, *fill = NULL, *rowpos = NULL;
    if (!probs) goto done;
    rc = fmo_gate(x, wg_eff, S, H, E, k, logits_out, probs, gate_out_out, topk_idx_out, mcw_out, abs_sum_out);
    if (rc) goto done;
    rc = fmo_slots(topk_idx_out, S, E, k, EC, slot_out, kept_out, counts_out);
    if (rc) goto done;
    rc = -2;
    yoff = (int64_t*)calloc((size_t)E + 1, sizeof(int64_t));
    fill = (int64_t*)calloc((size_t)E + 1, sizeof(int64_t));
    rowpos = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_sample ? n_sample : 1) * k);
    if (!yoff || !fill || !rowpos) goto done;
    for (int i = 0; i < n_sample; ++i) {
        const int t = sample[i];
        if (t < 0 || t >= S) { rc = -1; goto done; }
        for (int j = 0; j < k; ++j)
            if (kept_out[(size_t)t * k + j]) yoff[topk_idx_out[(size_t)t * k + j] + 1] += 1;
    }
    for (int e = 0; e < E; ++e) yoff[e + 1] += yoff[e];
    {
        const int64_t R = yoff[E];
        rows = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(R ? R : 1) * H);
        hbuf = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(R ? R : 1) * P);
        ybuf = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(R ? R : 1) * H);
        if (!rows || !hbuf || !ybuf) goto done;
        for (int i = 0; i < n_sample; ++i) {
            const int t = sample[i];
            for (int j = 0; j < k; ++j) {
                rowpos[(size_t)i * k + j] = -1;
                if (kept_out[(size_t)t * k + j]) {
                    const int e = topk_idx_out[(size_t)t * k + j];
                    const int64_t r = yoff[e] + fill[e]++;
                    rowpos[(size_t)i * k + j] = r;
                    memcpy(rows + (size_t)r * H, x + (size_t)t * H, sizeof(uint16_t) * (size_t)H);
                }
            }
        }
        for (int e = 0; e < E; ++e) {
            const int64_t r = yoff[e + 1] - yoff[e];
            rc = fmo_expert_ffn(rows + (size_t)yoff[e] * H, r, H, P, w_up + (size_t)e * P * H,
                                w_down_eff + (size_t)e * H * P, b_up ? b_up + (size_t)e * P : NULL,
                                b_down ? b_down + (size_t)e * H : NULL, act, hbuf + (size_t)yoff[e] * P,
                                ybuf + (size_t)yoff[e] * H);
            if (rc) goto done;
        }
    }
    for (int i = 0; i < n_sample; ++i) {
        const int t = sample[i];
        const uint16_t* yr[16];
        uint16_t pt[16];
        for (int j = 0; j < k; ++j) {
            const int e = topk_idx_out[(size_t)t * k + j];
            yr[j] = rowpos[(size_t)i * k + j] >= 0 ? ybuf + (size_t)rowpos[(size_t)i * k + j] * H : NULL;
            pt[j] = topk_w_given ? topk_w_given[(size_t)i * k + j] : gate_out_out[(size_t)t * E + e];
        }
        fmo_combine_token(yr, pt, mcw_given ? mcw_given[i] : mcw_out[t], k, H, out_sample + (size_t)i * H);
    }
    rc = 0;
done:
    free(rows); free(hbuf); free(ybuf); free(yoff); free(fill); free(rowpos); free(probs);
    return rc;
}

FMO_API int fmo_forward_sample(const uint16_t* x, const uint16_t* wg_eff, const uint16_t* w_up,
                               const uint16_t* w_down_eff, const uint16_t* b_up, const uint16_t* b_down, int S, int H,
                               int P, int E, int k, int EC, int act, const int32_t* sample, int n_sample,
                               uint16_t* out_sample, int32_t* topk_idx_out, int32_t* slot_out, int32_t* kept_out,
                               int32_t* counts_out, float* mcw_out, uint16_t* gate_out_out, float* logits_out,
                               float* abs_sum_out) {
    return fmo_forward_sample_w(x, wg_eff, w_up, w_down_eff, b_up, b_down, S, H, P, E, k, EC, act, sample, n_sample, NULL,
                                NULL, out_sample, topk_idx_out, slot_out, kept_out, counts_out, mcw_out, gate_out_out,
                                logits_out, abs_sum_out);
}

/* ------------------------------------------------------------------ training-mode auxiliary (load-balancing) loss
 * JobType::training only (is_training = 1).  The reference accumulates, per launch (buffers cleared by clearState,
 * moe/moe.cuh:49-54; layout gML[E] | gMeC[E] | gL, types.cuh:936-958):
 *   gML[e]  = sum over 128-token gate tiles of (column sum of the fp32 softmax probabilities of the tile) / S
 *             (moe/gate.cuh:608-635: BlockReduce per column, atomicAdd(gML + e, colAgg / S))  = mean_t p[t,e]
 *   gMeC[e] = sum over tiles of (selections of e in the tile) / S   (moe/gate.cuh:698-706)  = counts[e] / S, counting
 *             every selection, dropped or not
 *   gL      = sum_e gML[e] * gMeC[e] / E                              (moe/gate.cuh:763-773)
 * fp32 atomics in arrival order in the reference; double accumulation here (compare with a relative tolerance).
 * probs f32 [S,E] (fmo_gate's output), counts i32 [E] (fmo_slots' output).
 */
FMO_API void fmo_aux_loss(const float* probs, const int32_t* counts, int S, int E, float* gML, float* gMeC, float* loss) {
    double l = 0.0;
    for (int e = 0; e < E; ++e) {
        double acc = 0.0;
        for (int t = 0; t < S; ++t) acc += (double)probs[(size_t)t * E + e];
        gML[e] = (float)(acc / (double)S);
        gMeC[e] = (float)((double)counts[e] / (double)S);
        l += (double)gML[e] * (double)gMeC[e] / (double)E;
    }
    *loss = (float)l;
}

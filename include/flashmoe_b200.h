/*
 * flashmoe_b200.h -- C-ABI of the B200-native fused distributed-MoE forward path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  It replaces what the reference binds
 * through pybind in csrc/python_bindings.cu (reference paths below are relative to osayamenja/FlashMoE):
 *
 *   reference interface                                   replaced by
 *   ---------------------------------------------------   ------------------------------------------------------
 *   _C.initialize()        python_bindings.cu:157-159     fm_create + fm_symm_* (explicit rank/world, no NVSHMEM PMI)
 *                          bootstrap.cuh:532-547
 *   _C.finalize()          python_bindings.cu:164-166     fm_destroy
 *                          bootstrap.cuh:561-588
 *   _C.moe_forward(...)    python_bindings.cu:17-151      fm_moe_forward (device buffers, caller's weights used in
 *                          moe.cuh:146-205                place) / fm_moe_forward_host (host activations in/out)
 *   _C.get_compiled_config python_bindings.cu:170-179     fm_compiled_config + fm_get_dims
 *   _C.get_bookkeeping /   python_bindings.cu:181-189     fm_num_local_experts
 *   get_num_local_experts
 *   csrc/flashmoe_config.json -> -D macros -> ACC         fm_config_t (the same 15 keys); the JSON is baked into the
 *                          CMakeLists.txt:114-237         library at build time and returned by fm_compiled_config
 *                          types.cuh:441-512
 *
 * Error convention: every function returns 0 on success or a negative FM_E* code; fm_last_error() returns a
 * thread-local human-readable message.  The library never calls exit() (the reference does: debug.cuh:19-43).
 *
 * Threading: a context is used by one host thread at a time; one context (= one process) per GPU; work is enqueued
 * on the stream passed by the caller (the reference uses cudaStreamPerThread and blocks, python_bindings.cu:145).
 */
#ifndef FLASHMOE_B200_H
#define FLASHMOE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_API __attribute__((visibility("default")))

#define FM_OK 0
#define FM_EINVAL (-1)   /* bad argument / config violates a hot-path constraint */
#define FM_ECUDA (-2)    /* a CUDA runtime/driver call failed */
#define FM_ESTATE (-3)   /* call made in the wrong state (e.g. forward before peers are attached) */
#define FM_EKERNEL (-4)  /* the kernel reported a protocol timeout (see fm_last_error) */
#define FM_ENOMEM (-5)

#define FM_MAX_WORLD 16
#define FM_IPC_HANDLE_BYTES 64

/* The 15 keys of csrc/flashmoe_config.json (reference csrc/flashmoe_config.json:1-16), same names. */
typedef struct fm_config {
    int32_t capacity_factor;
    int32_t drop_tokens;
    int32_t expert_top_k;
    int32_t global_batch;
    int32_t is_training;
    int32_t hidden_act; /* 0 relu, 1 gelu */
    int32_t hidden_size;
    int32_t intermediate_size;
    int32_t mini_batch;
    int32_t moe_frequency;
    int32_t num_experts;
    int32_t num_layers;
    int32_t sequence_len;
    int32_t torch_dtype; /* 2 = bf16 is the only compute type of this build */
    int32_t vocab_size;
} fm_config_t;

/* Derived dimensions, the reference's get_compiled_config() keys plus the capacity figures (types.cuh:470-504). */
typedef struct fm_dims {
    int32_t S, H, E, P, PX, element_size;
    int32_t k, EC, pEC, TCM;
    int32_t world, rank, num_local_experts;
    int32_t num_sms, smem_bytes;
    int32_t grid; /* CTAs of the persistent kernel (= num_sms, rounded down to even for CTA pairs); rows of FM_BUF_TRACE */
} fm_dims_t;

typedef struct fm_ctx fm_ctx_t;

/* Config baked into the library at build time from csrc/flashmoe_config.json. */
FM_API int fm_compiled_config(fm_config_t* out);

/* Create a context on CUDA device `device` for rank `rank` of `world` expert-parallel ranks.
 * cfg == NULL uses the compiled config.  Allocates all workspaces, including this rank's symmetric slab
 * (receive buffers + flags that peers write into over NVLink). */
FM_API int fm_create(const fm_config_t* cfg, int rank, int world, int device, fm_ctx_t** out);
FM_API int fm_destroy(fm_ctx_t* ctx);

FM_API int fm_get_dims(const fm_ctx_t* ctx, fm_dims_t* out);
FM_API int fm_num_local_experts(const fm_ctx_t* ctx);

/* ---- symmetric memory plumbing (replaces nvshmem_malloc/nvshmem_ptr, bootstrap.cuh:359-360,442-443) ----
 * world == 1: nothing to do.  Otherwise the host exchanges either CUDA IPC handles (any out-of-band channel,
 * e.g. torch.distributed all_gather) or already-mapped peer base pointers (e.g. torch symmetric memory). */
FM_API int fm_symm_size(const fm_ctx_t* ctx, size_t* bytes);
FM_API int fm_symm_local_ptr(const fm_ctx_t* ctx, void** base);
FM_API int fm_symm_export(fm_ctx_t* ctx, void* handle_out /* FM_IPC_HANDLE_BYTES */);
FM_API int fm_symm_attach_ipc(fm_ctx_t* ctx, const void* all_handles /* world x FM_IPC_HANDLE_BYTES, rank order */);
FM_API int fm_symm_attach_ptrs(fm_ctx_t* ctx, void* const* peer_bases /* world pointers, rank order */);
/* Use a caller-provided slab (size >= fm_symm_size, 1 KiB aligned, zero-filled) instead of the internal one;
 * must be called before attach.  For torch symmetric memory / NVSHMEM allocations. */
FM_API int fm_symm_use_external(fm_ctx_t* ctx, void* base, size_t bytes);

/* ---- the hot path ----
 * All pointers are DEVICE pointers to contiguous bf16 tensors:
 *   x            [S, H]                     activations of this rank's tokens        (input.view(S,H))
 *   gate_w       [H, E]                     consumed REINTERPRETED flat as [E, H]    (python_bindings.cu:93-99)
 *   expert_w     [nLx, 2, P, H]             [i,0] = W_up [P,H]; [i,1] flat-viewed as W_down [H,P] (:104-119)
 *   bias_up      [nLx, P] or NULL (zero)    bias_down [nLx, H] or NULL (zero)        (moe.cuh:117-124)
 *   out          [S, H]                     written in full
 * Enqueues one fused persistent kernel on `stream` (cudaStream_t); collective across the world: every rank must
 * call it the same number of times.  Does not synchronise. */
FM_API int fm_moe_forward(fm_ctx_t* ctx, const void* x, const void* gate_w, const void* expert_w,
                          const void* bias_up, const void* bias_down, void* out, void* stream);

/* world > 1: the [S,H] bf16 region of this rank's symmetric slab into which the experts (local and remote) accumulate
 * this rank's output rows.  Passing exactly this pointer as `out` to fm_moe_forward leaves the result there and skips
 * the final copy into a caller tensor (zero-copy output, like registering a user buffer with NVSHMEM); it is
 * overwritten by the next forward.  world == 1: *ptr = NULL, *bytes = 0 (the kernel accumulates straight into `out`). */
FM_API int fm_output_buffer(const fm_ctx_t* ctx, void** ptr, size_t* bytes);

/* Same, with the activations in HOST memory (pinned for full PCIe speed): copies x host->device, runs the layer,
 * copies out device->host, then waits for the result (the reference's blocking behaviour,
 * python_bindings.cu:131-148).  Equivalent to fm_host_submit + fm_host_wait. */
FM_API int fm_moe_forward_host(fm_ctx_t* ctx, const void* x_host, const void* gate_w, const void* expert_w,
                               const void* bias_up, const void* bias_down, void* out_host, void* stream);

/* Pipelined form of the host-buffer call: fm_host_submit enqueues H2D copy -> layer -> D2H copy for one step and
 * returns a ticket at once; up to FM_HOST_SLOTS steps may be in flight, so the copies of neighbouring steps overlap
 * the kernel (separate copy streams, PCIe full duplex).  The kernel runs on `stream`; weights must be ready there.
 * fm_host_wait blocks until that step's out_host is complete; tickets are waited for in submission order.
 * x_host must stay valid until its H2D copy has run (i.e. until the step's ticket has been waited for). */
#define FM_HOST_SLOTS 3
FM_API int fm_host_submit(fm_ctx_t* ctx, const void* x_host, const void* gate_w, const void* expert_w,
                          const void* bias_up, const void* bias_down, void* out_host, void* stream, uint64_t* ticket);
FM_API int fm_host_wait(fm_ctx_t* ctx, uint64_t ticket);

/* Check the kernel's host-mapped status record after a synchronise; FM_EKERNEL + message if a wait timed out. */
FM_API int fm_check(fm_ctx_t* ctx);
FM_API int fm_set_timeout_ms(fm_ctx_t* ctx, uint32_t ms);
FM_API uint64_t fm_launch_count(const fm_ctx_t* ctx);
/* Per-CTA phase / tile time stamps into FM_BUF_TRACE (profiling aid; off by default). */
FM_API int fm_set_trace(fm_ctx_t* ctx, int enable);

/* ---- inspection (tests / debugging): copy an internal device buffer to host memory after synchronising ---- */
enum fm_buffer {
    FM_BUF_TOPK_IDX = 0, /* int32 [S,k]   expert index per pick (bit-exact parity target) */
    FM_BUF_TOPK_W = 1,   /* bf16  [S,k]   gateOut[t, e_j] (bf16-rounded probability) */
    FM_BUF_MCW = 2,      /* f32   [S]     sum of the k selected fp32 probabilities */
    FM_BUF_SLOT = 3,     /* int32 [S,k]   slot in the (this rank, expert) packet; >= EC means dropped */
    FM_BUF_COUNTS = 4,   /* int32 [E]     selections per expert on this rank (incl. dropped) */
    FM_BUF_RECV_X = 5,   /* bf16  [W, nLx, pEC, H]  dispatched rows received by this rank */
    FM_BUF_HIDDEN = 6,   /* bf16  [W, nLx, pEC, P]  h = act(x W_up^T + b) staging */
    FM_BUF_RET_Y = 7,    /* bf16  [E, pEC, H]       expert outputs returned to this rank */
    FM_BUF_GATE_OUT = 8, /* bf16  [S, E]  full softmax row (reference gateOut[S,PX] without the padding columns) */
    FM_BUF_RECV_CNT = 9, /* int32 [W, nLx] rows received per (source rank, local expert) in the last forward */
    FM_BUF_TRACE = 10,   /* u64   [grid, 128] %globaltimer stamps of the last forward (fm_set_trace) */
    FM_BUF_AUX_LOSS = 11 /* f32   [2E+1] is_training = 1 only: gML[E] (mean gate probability per expert), gMeC[E] (mean
                            routed fraction per expert), loss = sum_e gML*gMeC / E  (moe/gate.cuh:608-635,698-706,763-773) */
};
FM_API int fm_buffer_bytes(const fm_ctx_t* ctx, int which, size_t* bytes);
FM_API int fm_read_buffer(fm_ctx_t* ctx, int which, void* host_dst, size_t bytes);

/* Debug: run only the phases in `phase_mask` (bit0 gate+dispatch, bit1 expert FFN, bit2 combine).  A launch that
 * omits bit0 reuses the previous launch's routing and epoch (single rank only). */
FM_API int fm_debug_forward(fm_ctx_t* ctx, const void* x, const void* gate_w, const void* expert_w,
                            const void* bias_up, const void* bias_down, void* out, void* stream,
                            uint32_t phase_mask);

FM_API const char* fm_last_error(void);
FM_API const char* fm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FLASHMOE_B200_H */

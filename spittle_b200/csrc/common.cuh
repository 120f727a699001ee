// Shared helpers for the spittle_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <atomic>
#include <cstring>

#include "../../include/spittle_b200.h"

namespace sb {

// thread-local last error (sb_last_error)
void set_error(const std::string& msg);
const char* get_error();

#define SB_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::sb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) +         \
                            " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");     \
            return SB_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define SB_CHECK_ARG(cond, msg)                                                          \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            ::sb::set_error(std::string("invalid argument: ") + (msg));                  \
            return SB_ERR_INVALID;                                                       \
        }                                                                                \
    } while (0)

// 16-bit operand type traits: the engine runs either bf16 (north-star dtype) or f16
// (the reference's own rounding points: ggml f16 weights x f16-rounded activations).
template <typename T> struct Op16;
template <> struct Op16<__nv_bfloat16> {
    static constexpr int kUmmaFormat = 1;  // UMMA F16F32Format::BF16
    __device__ __forceinline__ static __nv_bfloat16 from_f32(float x) { return __float2bfloat16_rn(x); }
    __device__ __forceinline__ static float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
    __device__ __forceinline__ static uint32_t pack2(float a, float b) {
        __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __device__ __forceinline__ static float2 unpack2(uint32_t u) {
        __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
        return __bfloat1622float2(v);
    }
};
template <> struct Op16<__half> {
    static constexpr int kUmmaFormat = 0;  // UMMA F16F32Format::F16
    __device__ __forceinline__ static __half from_f32(float x) { return __float2half_rn(x); }
    __device__ __forceinline__ static float to_f32(__half x) { return __half2float(x); }
    __device__ __forceinline__ static uint32_t pack2(float a, float b) {
        __half2 v = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __device__ __forceinline__ static float2 unpack2(uint32_t u) {
        __half2 v = *reinterpret_cast<__half2*>(&u);
        return __half22float2(v);
    }
};

__device__ __forceinline__ float gelu_tanh(float x) {
    // 0.5 x (1 + tanh(sqrt(2/pi) x (1 + 0.044715 x^2)))   (whisper.cpp / ggml GELU, App. C.2)
    // 0.5 (1 + tanh u) == 1 / (1 + exp(-2u)): two MUFU ops (ex2.approx, rcp.approx) + 5 FMA-pipe ops
    // instead of tanhf's ~20 instructions; relative error ~1e-6, far below the 16-bit rounding that follows
    const float k2 = -2.0f * 0.79788456080286535588f * 1.4426950408889634f;   // -2 sqrt(2/pi) log2(e)
    const float t = x * fmaf(0.044715f * x, x, 1.0f);
    return __fdividef(x, 1.0f + exp2f(k2 * t));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// monotone float <-> int key (for atomicMax on floats of either sign)
__host__ __device__ __forceinline__ int float_key(float f) {
#ifdef __CUDA_ARCH__
    int i = __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4);
#endif
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float key_float(int k) {
    int i = k >= 0 ? k : k ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Decoder-step kernels are ~140 small dependent launches per token.  Each calls pdl_wait() before
// touching anything a predecessor writes and pdl_trigger() once its main work is issued, so the next
// kernel is scheduled and runs its independent prologue (weight prefetch, first K tile) under this
// kernel's tail.  Triggering at the very top was measured slower: early-resident CTAs of the following
// kernels steal registers/smem from the bandwidth-bound cross-attention.  Data produced by a
// predecessor is always read through L2 (__ldcg): co-resident kernels can leave stale lines in L1,
// which is only invalidated at launch.  Without the launch attribute both calls are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Device-side launch trace (sb_engine_set_profile(e, 2)): thread 0 of every block stamps %globaltimer into the slot
// of its launch (slot = lane step x launches-per-step + index of the launch inside the step; min over the
// blocks for the start, max for the end).  The launches keep their CUDA graph and their PDL overlap, so these are the
// durations inside the real chain, which CUDA-event brackets around single launches cannot give.
struct TraceSlot {
    unsigned long long* t0 = nullptr;   // [max_steps * per_step] first block start (ns)
    unsigned long long* t1 = nullptr;   // [max_steps * per_step] last block end (ns)
    const int* tick = nullptr;          // device: step counter of the lane
    int idx = 0, per_step = 0, max_steps = 0;
};
#ifdef __CUDACC__
__device__ __forceinline__ void trace_begin(const TraceSlot& ts) {
    if (ts.t0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const int step = __ldcg(ts.tick);
        if (step < ts.max_steps) atomicMin(ts.t0 + (size_t)step * ts.per_step + ts.idx, t);
    }
}
__device__ __forceinline__ void trace_end(const TraceSlot& ts) {
    if (ts.t1 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const int step = __ldcg(ts.tick);
        if (step < ts.max_steps) atomicMax(ts.t1 + (size_t)step * ts.per_step + ts.idx, t);
    }
}
#endif
extern thread_local TraceSlot g_trace_next;   // core.cu: consumed (and cleared) by the next decoder-stage launcher

extern bool g_pdl;   // core.cu; env SB_PDL=0 disables

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// cudaFuncSetAttribute is PER DEVICE and one process may drive several GPUs (sb_config.devices): `SB_ONCE_PER_DEVICE(stmt)`
// runs `stmt` the first time this call site is reached with each device current (set twice under a race: harmless).
#define SB_ONCE_PER_DEVICE(...)                                                          \
    do {                                                                                 \
        static std::atomic<uint64_t> _sb_mask{0};                                        \
        int _sb_dev = 0;                                                                 \
        cudaGetDevice(&_sb_dev);                                                         \
        const uint64_t _sb_bit = 1ull << (_sb_dev & 63);                                 \
        if (!(_sb_mask.load(std::memory_order_acquire) & _sb_bit)) {                     \
            __VA_ARGS__;                                                                  \
            _sb_mask.fetch_or(_sb_bit, std::memory_order_release);                       \
        }                                                                                \
    } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace sb

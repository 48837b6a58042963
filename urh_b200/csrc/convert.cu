// Sample-format conversions of IQArray.convert_to (IQArray.py:127-200) on the GPU: the capture formats
// cs8 / cu8 / cs16 / cu16 / float32 into each other, element by element, with numpy's integer wrap-around and C's
// float -> int truncation.  One pass, 1..4 bytes read and written per element.
#include "common.cuh"

#include <type_traits>

// numpy astype(float32 -> small int) is a C cast: on x86-64 cvttss2si to int32 (INT_MIN for NaN / out of range), low bits kept
__device__ __forceinline__ int32_t c_cast_i32(float v) {
    if (!(fabsf(v) < 2147483648.0f)) return (int32_t)0x80000000;
    return __float2int_rz(v);
}

template <typename S, typename D>
__device__ __forceinline__ D conv_one(S x);

// ---- from uint8 (IQArray.py:131-143)
template <> __device__ __forceinline__ int8_t conv_one<uint8_t, int8_t>(uint8_t x) { return (int8_t)(uint8_t)(x - 128u); }
template <> __device__ __forceinline__ int16_t conv_one<uint8_t, int16_t>(uint8_t x) { return (int16_t)(uint16_t)(((int)x - 128) << 8); }
template <> __device__ __forceinline__ uint16_t conv_one<uint8_t, uint16_t>(uint8_t x) { return (uint16_t)((unsigned)x << 8); }
template <> __device__ __forceinline__ float conv_one<uint8_t, float>(uint8_t x) { return __fadd_rn(__fmul_rn((float)x, 0.0078125f), -1.0f); }
// ---- from int8 (:145-153)
template <> __device__ __forceinline__ uint8_t conv_one<int8_t, uint8_t>(int8_t x) { return (uint8_t)((int)x + 128); }
template <> __device__ __forceinline__ int16_t conv_one<int8_t, int16_t>(int8_t x) { return (int16_t)(uint16_t)((int)x << 8); }
template <> __device__ __forceinline__ uint16_t conv_one<int8_t, uint16_t>(int8_t x) { return (uint16_t)(((int)x + 128) << 8); }
template <> __device__ __forceinline__ float conv_one<int8_t, float>(int8_t x) { return __fmul_rn((float)x, 0.0078125f); }
// ---- from uint16 (:155-170)
template <> __device__ __forceinline__ int8_t conv_one<uint16_t, int8_t>(uint16_t x) { return (int8_t)(((int16_t)(uint16_t)(x - 32768u)) >> 8); }
template <> __device__ __forceinline__ uint8_t conv_one<uint16_t, uint8_t>(uint16_t x) { return (uint8_t)(x >> 8); }
template <> __device__ __forceinline__ int16_t conv_one<uint16_t, int16_t>(uint16_t x) { return (int16_t)(uint16_t)(x - 32768u); }
template <> __device__ __forceinline__ float conv_one<uint16_t, float>(uint16_t x) { return __fadd_rn(__fmul_rn((float)x, 3.0517578125e-05f), -1.0f); }
// ---- from int16 (:172-183)
template <> __device__ __forceinline__ int8_t conv_one<int16_t, int8_t>(int16_t x) { return (int8_t)(x >> 8); }
template <> __device__ __forceinline__ uint8_t conv_one<int16_t, uint8_t>(int16_t x) { return (uint8_t)(((uint16_t)((int)x + 32768)) >> 8); }
template <> __device__ __forceinline__ uint16_t conv_one<int16_t, uint16_t>(int16_t x) { return (uint16_t)((int)x + 32768); }
template <> __device__ __forceinline__ float conv_one<int16_t, float>(int16_t x) { return __fmul_rn((float)x, 3.0517578125e-05f); }
// ---- from float32 (:185-200)
template <> __device__ __forceinline__ int8_t conv_one<float, int8_t>(float x) { return (int8_t)c_cast_i32(__fmul_rn(x, 127.0f)); }
template <> __device__ __forceinline__ uint8_t conv_one<float, uint8_t>(float x) { return (uint8_t)c_cast_i32(__fmul_rn(__fadd_rn(x, 1.0f), 127.0f)); }
template <> __device__ __forceinline__ int16_t conv_one<float, int16_t>(float x) { return (int16_t)c_cast_i32(__fmul_rn(x, 32767.0f)); }
template <> __device__ __forceinline__ uint16_t conv_one<float, uint16_t>(float x) { return (uint16_t)c_cast_i32(__fmul_rn(__fadd_rn(x, 1.0f), 32767.0f)); }

template <typename S, typename D>
__global__ void k_convert(const S* __restrict__ in, D* __restrict__ out, int64_t count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) out[i] = conv_one<S, D>(in[i]);
}

template <typename S>
static int convert_from(urh_ctx* ctx, const void* d_in, void* d_out, int out_dtype, int64_t count, unsigned grid) {
    switch (out_dtype) {
        case URH_DT_I8: if constexpr (!std::is_same<S, int8_t>::value) { URH_LAUNCH(ctx, (k_convert<S, int8_t>), grid, 256, 0, (const S*)d_in, (int8_t*)d_out, count); return URH_OK; } break;
        case URH_DT_U8: if constexpr (!std::is_same<S, uint8_t>::value) { URH_LAUNCH(ctx, (k_convert<S, uint8_t>), grid, 256, 0, (const S*)d_in, (uint8_t*)d_out, count); return URH_OK; } break;
        case URH_DT_I16: if constexpr (!std::is_same<S, int16_t>::value) { URH_LAUNCH(ctx, (k_convert<S, int16_t>), grid, 256, 0, (const S*)d_in, (int16_t*)d_out, count); return URH_OK; } break;
        case URH_DT_U16: if constexpr (!std::is_same<S, uint16_t>::value) { URH_LAUNCH(ctx, (k_convert<S, uint16_t>), grid, 256, 0, (const S*)d_in, (uint16_t*)d_out, count); return URH_OK; } break;
        case URH_DT_F32: if constexpr (!std::is_same<S, float>::value) { URH_LAUNCH(ctx, (k_convert<S, float>), grid, 256, 0, (const S*)d_in, (float*)d_out, count); return URH_OK; } break;
        default: URH_FAIL(ctx, URH_ERR_DTYPE, "Data type not supported");
    }
    // same type: plain copy
    URH_CUDA(ctx, cudaMemcpyAsync(d_out, d_in, (size_t)count * sizeof(S), cudaMemcpyDeviceToDevice, ctx->stream));
    return URH_OK;
}

// count = number of ELEMENTS (2 per IQ sample).  Asynchronous on the context's stream.
extern "C" int urh_convert_iq(urh_ctx* ctx, const void* d_in, int in_dtype, void* d_out, int out_dtype, int64_t count) {
    if (count <= 0) return URH_OK;
    const unsigned grid = (unsigned)min(urh_div_up(count, 256), (int64_t)ctx->sm_count * 32);
    switch (in_dtype) {
        case URH_DT_I8: return convert_from<int8_t>(ctx, d_in, d_out, out_dtype, count, grid);
        case URH_DT_U8: return convert_from<uint8_t>(ctx, d_in, d_out, out_dtype, count, grid);
        case URH_DT_I16: return convert_from<int16_t>(ctx, d_in, d_out, out_dtype, count, grid);
        case URH_DT_U16: return convert_from<uint16_t>(ctx, d_in, d_out, out_dtype, count, grid);
        case URH_DT_F32: return convert_from<float>(ctx, d_in, d_out, out_dtype, count, grid);
        default: URH_FAIL(ctx, URH_ERR_DTYPE, "Data type not supported");
    }
}

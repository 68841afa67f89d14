// LSTM cell kernels (fp32) for the recurrent actor-critic (BASELINE config 4).
//
// Replaces flax nn.OptimizedLSTMCell as used by MultiLayerLSTMCell / LSTM (ml/rnn.py:10-111):
// gates i, f, g, o; z = x W_i* + h W_h* + b_h*; c' = sigmoid(f) c + sigmoid(i) tanh(g);
// h' = sigmoid(o) tanh(c'); LSTM.sequence zeroes the carry AFTER a step whose `end` flag is
// set (ml/rnn.py:91-96, clear_recurrent_state :66-81) while the step's output stays unmasked.
// The matrix products (input projection for all time steps at once, the recurrent h W_h per
// step, and their transposes in BPTT) are the GEMM entry points; these kernels are the
// element-wise cell math around them, 128-bit vectorised, one thread per 4 hidden units.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// z [M, 4H] (gate-major blocks i|f|g|o), bias [4H], c_prev [M, H]
// h_seq [M, H]: unmasked output;  c_carry/h_carry [M, H]: state for the next step (masked by ends)
// stash (may be NULL) [M, 5H]: i, f, g, o, tanh(c')
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = o;
}

// TH: type of the unmasked output h_seq (float, or bf16 on the tensor-core path, where it is the A operand
// of the head GEMM); h_carry_bf (tensor-core path, may be NULL): bf16 copy of the masked carry = the A
// operand of the next step's recurrent GEMM.  bias may be NULL (already added by the GEMM epilogue).
template <typename TH>
__global__ void __launch_bounds__(256)
lstm_cell_fwd_kernel(const float* __restrict__ z, const float* __restrict__ bias,
                     const float* __restrict__ c_prev, const uint8_t* __restrict__ ends,
                     TH* __restrict__ h_seq, float* __restrict__ c_carry, float* __restrict__ h_carry,
                     __nv_bfloat16* __restrict__ h_carry_bf, float* __restrict__ stash, long long M, int H) {
    const int hv = H / 4;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * hv) return;
    const long long m = t / hv;
    const int j = (int)(t - m * hv) * 4;
    const float* zr = z + m * 4 * H;
    const float4 zi = *reinterpret_cast<const float4*>(zr + j);
    const float4 zf = *reinterpret_cast<const float4*>(zr + H + j);
    const float4 zg = *reinterpret_cast<const float4*>(zr + 2 * H + j);
    const float4 zo = *reinterpret_cast<const float4*>(zr + 3 * H + j);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 bi = bias ? *reinterpret_cast<const float4*>(bias + j) : z4;
    const float4 bf = bias ? *reinterpret_cast<const float4*>(bias + H + j) : z4;
    const float4 bg = bias ? *reinterpret_cast<const float4*>(bias + 2 * H + j) : z4;
    const float4 bo = bias ? *reinterpret_cast<const float4*>(bias + 3 * H + j) : z4;
    const float4 cp = *reinterpret_cast<const float4*>(c_prev + m * H + j);
    const float keep = (ends && ends[m]) ? 0.f : 1.f;
    float i4[4], f4[4], g4[4], o4[4], tc4[4], c4[4], h4[4];
    const float zi_[4] = {zi.x + bi.x, zi.y + bi.y, zi.z + bi.z, zi.w + bi.w};
    const float zf_[4] = {zf.x + bf.x, zf.y + bf.y, zf.z + bf.z, zf.w + bf.w};
    const float zg_[4] = {zg.x + bg.x, zg.y + bg.y, zg.z + bg.z, zg.w + bg.w};
    const float zo_[4] = {zo.x + bo.x, zo.y + bo.y, zo.z + bo.z, zo.w + bo.w};
    const float cp_[4] = {cp.x, cp.y, cp.z, cp.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        i4[k] = sigmoidf_(zi_[k]); f4[k] = sigmoidf_(zf_[k]); g4[k] = tanhf(zg_[k]); o4[k] = sigmoidf_(zo_[k]);
        c4[k] = f4[k] * cp_[k] + i4[k] * g4[k];
        tc4[k] = tanhf(c4[k]);
        h4[k] = o4[k] * tc4[k];
    }
    store4(h_seq + m * H + j, h4);
    *reinterpret_cast<float4*>(c_carry + m * H + j) = make_float4(keep * c4[0], keep * c4[1], keep * c4[2], keep * c4[3]);
    const float hk[4] = {keep * h4[0], keep * h4[1], keep * h4[2], keep * h4[3]};
    if (h_carry) store4(h_carry + m * H + j, hk);
    if (h_carry_bf) store4(h_carry_bf + m * H + j, hk);
    if (stash) {
        float* s = stash + m * 5 * H + j;
        *reinterpret_cast<float4*>(s) = make_float4(i4[0], i4[1], i4[2], i4[3]);
        *reinterpret_cast<float4*>(s + H) = make_float4(f4[0], f4[1], f4[2], f4[3]);
        *reinterpret_cast<float4*>(s + 2 * H) = make_float4(g4[0], g4[1], g4[2], g4[3]);
        *reinterpret_cast<float4*>(s + 3 * H) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        *reinterpret_cast<float4*>(s + 4 * H) = make_float4(tc4[0], tc4[1], tc4[2], tc4[3]);
    }
}

// BPTT through one cell step.  dh_seq [M, ld_dh]: gradient w.r.t. the unmasked output of this step;
// dh_carry/dc_carry [M, H]: gradients w.r.t. the (masked) carry leaving this step (NULL at the
// last step); -> dz [M, 4H], dc_prev [M, H]  (dh_prev = dz W_h is a GEMM done by the caller)
template <typename TZ>
__global__ void __launch_bounds__(256)
lstm_cell_bwd_kernel(const float* __restrict__ dh_seq, int ld_dh, const float* __restrict__ dh_carry,
                     const float* __restrict__ dc_carry, const uint8_t* __restrict__ ends,
                     const float* __restrict__ stash, const float* __restrict__ c_prev,
                     TZ* __restrict__ dz, float* __restrict__ dc_prev, long long M, int H) {
    const int hv = H / 4;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * hv) return;
    const long long m = t / hv;
    const int j = (int)(t - m * hv) * 4;
    const float keep = (ends && ends[m]) ? 0.f : 1.f;
    const float* s = stash + m * 5 * H + j;
    const float4 i4 = *reinterpret_cast<const float4*>(s);
    const float4 f4 = *reinterpret_cast<const float4*>(s + H);
    const float4 g4 = *reinterpret_cast<const float4*>(s + 2 * H);
    const float4 o4 = *reinterpret_cast<const float4*>(s + 3 * H);
    const float4 t4 = *reinterpret_cast<const float4*>(s + 4 * H);
    const float4 cp = *reinterpret_cast<const float4*>(c_prev + m * H + j);
    float4 dh = *reinterpret_cast<const float4*>(dh_seq + m * ld_dh + j);
    float4 dc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dh_carry) {
        const float4 a = *reinterpret_cast<const float4*>(dh_carry + m * H + j);
        const float4 b = *reinterpret_cast<const float4*>(dc_carry + m * H + j);
        dh.x += keep * a.x; dh.y += keep * a.y; dh.z += keep * a.z; dh.w += keep * a.w;
        dc = make_float4(keep * b.x, keep * b.y, keep * b.z, keep * b.w);
    }
    const float i_[4] = {i4.x, i4.y, i4.z, i4.w}, f_[4] = {f4.x, f4.y, f4.z, f4.w};
    const float g_[4] = {g4.x, g4.y, g4.z, g4.w}, o_[4] = {o4.x, o4.y, o4.z, o4.w};
    const float t_[4] = {t4.x, t4.y, t4.z, t4.w}, cp_[4] = {cp.x, cp.y, cp.z, cp.w};
    const float dh_[4] = {dh.x, dh.y, dh.z, dh.w};
    float dc_[4] = {dc.x, dc.y, dc.z, dc.w};
    float dzi[4], dzf[4], dzg[4], dzo[4], dcp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float d_o = dh_[k] * t_[k];
        dc_[k] += dh_[k] * o_[k] * (1.f - t_[k] * t_[k]);
        dzi[k] = dc_[k] * g_[k] * i_[k] * (1.f - i_[k]);
        dzf[k] = dc_[k] * cp_[k] * f_[k] * (1.f - f_[k]);
        dzg[k] = dc_[k] * i_[k] * (1.f - g_[k] * g_[k]);
        dzo[k] = d_o * o_[k] * (1.f - o_[k]);
        dcp[k] = dc_[k] * f_[k];
    }
    TZ* zr = dz + m * 4 * H + j;
    store4(zr, dzi);
    store4(zr + H, dzf);
    store4(zr + 2 * H, dzg);
    store4(zr + 3 * H, dzo);
    *reinterpret_cast<float4*>(dc_prev + m * H + j) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
}

// rnn_reset_fn / clear_recurrent_state (ml/rnn.py:66-81): state[m, :] = 0 where dones[m]
__global__ void __launch_bounds__(256)
rnn_reset_kernel(float* __restrict__ state, const uint8_t* __restrict__ dones, long long M, int H) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M * H) return;
    if (dones[t / H]) state[t] = 0.f;
}

}  // namespace

MLB_API int mlb_lstm_cell_fwd_f32(void* stream, const float* z, const float* bias, const float* c_prev,
                                  const uint8_t* ends, float* h_seq, float* c_carry, float* h_carry,
                                  float* stash, long long M, int H) {
    MLB_REQUIRE(z && bias && c_prev && h_seq && c_carry && h_carry && M >= 0 && H > 0 && H % 4 == 0);
    if (M == 0) return MLB_OK;
    MLB_REQUIRE(mlb_aligned16(z) && mlb_aligned16(bias) && mlb_aligned16(c_prev) && mlb_aligned16(h_seq) &&
                mlb_aligned16(c_carry) && mlb_aligned16(h_carry) && (!stash || mlb_aligned16(stash)));
    lstm_cell_fwd_kernel<float><<<mlb_cdiv(M * (H / 4), 256), 256, 0, mlb_stream(stream)>>>(
        z, bias, c_prev, ends, h_seq, c_carry, h_carry, nullptr, stash, M, H);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_lstm_cell_bwd_f32(void* stream, const float* dh_seq, int ld_dh, const float* dh_carry,
                                  const float* dc_carry, const uint8_t* ends, const float* stash,
                                  const float* c_prev, float* dz, float* dc_prev, long long M, int H) {
    MLB_REQUIRE(dh_seq && stash && c_prev && dz && dc_prev && M >= 0 && H > 0 && H % 4 == 0 && ld_dh % 4 == 0);
    MLB_REQUIRE((dh_carry == nullptr) == (dc_carry == nullptr));
    if (M == 0) return MLB_OK;
    lstm_cell_bwd_kernel<float><<<mlb_cdiv(M * (H / 4), 256), 256, 0, mlb_stream(stream)>>>(
        dh_seq, ld_dh, dh_carry, dc_carry, ends, stash, c_prev, dz, dc_prev, M, H);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

// Tensor-core path (compute_dtype = bfloat16): the recurrent products run on tcgen05 (mlb_gemm_bf16_tc,
// fp32 accumulation into z), the cell state stays fp32; h_seq / h_carry_bf / dz are the bf16 GEMM operands.
MLB_API int mlb_lstm_cell_fwd_tc(void* stream, const float* z, const float* bias, const float* c_prev,
                                 const uint8_t* ends, void* h_seq_bf16, float* c_carry, float* h_carry,
                                 void* h_carry_bf16, float* stash, long long M, int H) {
    MLB_REQUIRE(z && c_prev && h_seq_bf16 && c_carry && M >= 0 && H > 0 && H % 4 == 0);
    if (M == 0) return MLB_OK;
    MLB_REQUIRE(mlb_aligned16(z) && mlb_aligned16(c_prev) && mlb_aligned16(c_carry) && (!bias || mlb_aligned16(bias)) &&
                (!h_carry || mlb_aligned16(h_carry)) && (!stash || mlb_aligned16(stash)) &&
                (reinterpret_cast<uintptr_t>(h_seq_bf16) & 7) == 0 && (reinterpret_cast<uintptr_t>(h_carry_bf16) & 7) == 0);
    lstm_cell_fwd_kernel<__nv_bfloat16><<<mlb_cdiv(M * (H / 4), 256), 256, 0, mlb_stream(stream)>>>(
        z, bias, c_prev, ends, static_cast<__nv_bfloat16*>(h_seq_bf16), c_carry, h_carry,
        static_cast<__nv_bfloat16*>(h_carry_bf16), stash, M, H);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_lstm_cell_bwd_tc(void* stream, const float* dh_seq, int ld_dh, const float* dh_carry,
                                 const float* dc_carry, const uint8_t* ends, const float* stash,
                                 const float* c_prev, void* dz_bf16, float* dc_prev, long long M, int H) {
    MLB_REQUIRE(dh_seq && stash && c_prev && dz_bf16 && dc_prev && M >= 0 && H > 0 && H % 4 == 0 && ld_dh % 4 == 0);
    MLB_REQUIRE((dh_carry == nullptr) == (dc_carry == nullptr) && (reinterpret_cast<uintptr_t>(dz_bf16) & 7) == 0);
    if (M == 0) return MLB_OK;
    lstm_cell_bwd_kernel<__nv_bfloat16><<<mlb_cdiv(M * (H / 4), 256), 256, 0, mlb_stream(stream)>>>(
        dh_seq, ld_dh, dh_carry, dc_carry, ends, stash, c_prev, static_cast<__nv_bfloat16*>(dz_bf16), dc_prev, M, H);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

MLB_API int mlb_rnn_reset_f32(void* stream, float* state, const uint8_t* dones, long long M, int H) {
    MLB_REQUIRE(state && dones && M >= 0 && H > 0);
    if (M == 0) return MLB_OK;
    rnn_reset_kernel<<<mlb_cdiv(M * H, 256), 256, 0, mlb_stream(stream)>>>(state, dones, M, H);
    MLB_CHECK_LAUNCH();
    return MLB_OK;
}

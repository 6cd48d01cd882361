// Tensor-core (tcgen05) STFT path — placeholder until the kernels land.
#include "brv_common.cuh"

int brv_tc_plan_init(brv_stft_plan* p, const std::vector<double>& fwd,
                     const std::vector<double>& inv) {
    (void)p; (void)fwd; (void)inv;
    return BRV_OK;
}
void brv_tc_plan_free(brv_stft_plan* p) { (void)p; }
bool brv_tc_supports_forward(const brv_stft_plan* p) { (void)p; return false; }
int brv_tc_stft_forward(const brv_stft_plan* p, const float* x, int64_t n_sig, int64_t samples,
                        int64_t x_stride, float2* out, int64_t n_frames, cudaStream_t st) {
    (void)p; (void)x; (void)n_sig; (void)samples; (void)x_stride; (void)out; (void)n_frames; (void)st;
    return brv_fail(BRV_ERR_UNSUPPORTED, "tensor-core path not built");
}

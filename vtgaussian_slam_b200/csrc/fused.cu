// fused.cu -- the per-iteration work the reference does in PyTorch around the rasteriser:
// the tracking / mapping loss of get_loss (reference src/vtgaussian_slam.py:513-612,678-679)
// with its gradient w.r.t. the rendered planes, and the Adam update
// (torch.optim.Adam as configured at src/vtgaussian_slam.py:180-187).
#include "common.cuh"
#include "kernels.h"

namespace vtgs {

constexpr int LOSS_TERMS = 4;      // per-block partials: depth L1, rgb L1, mask count, spare

__device__ __forceinline__ float sgn(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

// Tracking loss (mode 0): sums of masked absolute differences; dL/dplane = w * sign * mask.
__global__ void __launch_bounds__(256)
tracking_loss_kernel(const __grid_constant__ CamConst cam, VtgsLossConfig cfg, const float* __restrict__ image6,
                     const float* __restrict__ gt_rgb, const float* __restrict__ gt_depth,
                     float* __restrict__ dL_dimage4, float* __restrict__ partials) {
    __shared__ float s_part[8][LOSS_TERMS];
    const size_t P = (size_t)cam.W * cam.H;
    const size_t row_begin = (size_t)cam.row0 * 16 * cam.W;
    const size_t row_end = min(P, (size_t)cam.row1 * 16 * cam.W);
    const size_t pid = row_begin + (size_t)blockIdx.x * 256 + threadIdx.x;
    float ld = 0.f, li = 0.f, cnt = 0.f;
    if (pid < row_end) {
        const float r = image6[pid], g = image6[P + pid], b = image6[2 * P + pid];
        const float d = image6[3 * P + pid], sil = image6[4 * P + pid], dsq = image6[5 * P + pid];
        const float gd = gt_depth[pid];
        const float unc = dsq - d * d;
        bool mask = gd > 0.0f && !(d != d) && !(unc != unc);
        if (cfg.use_sil_for_loss) mask = mask && sil > cfg.sil_thres;
        if (cfg.far_depth_thres > 0.0f) mask = mask && gd < cfg.far_depth_thres;
        // reference :600-605: the colour term is masked only with use_sil_for_loss / outlier masks
        const bool mask_im = cfg.use_sil_for_loss ? mask : true;
        const float er = r - gt_rgb[pid], eg = g - gt_rgb[P + pid], eb = b - gt_rgb[2 * P + pid];
        const float ed = d - gd;
        if (mask) { ld = fabsf(ed); cnt = 1.0f; }
        if (mask_im) li = fabsf(er) + fabsf(eg) + fabsf(eb);
        dL_dimage4[pid] = mask_im ? cfg.w_im * sgn(er) : 0.0f;
        dL_dimage4[P + pid] = mask_im ? cfg.w_im * sgn(eg) : 0.0f;
        dL_dimage4[2 * P + pid] = mask_im ? cfg.w_im * sgn(eb) : 0.0f;
        dL_dimage4[3 * P + pid] = mask ? cfg.w_depth * sgn(ed) : 0.0f;
    }
    ld = warp_sum(ld); li = warp_sum(li); cnt = warp_sum(cnt);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_part[warp][0] = ld; s_part[warp][1] = li; s_part[warp][2] = cnt; s_part[warp][3] = 0.f; }
    __syncthreads();
    if (threadIdx.x < LOSS_TERMS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += s_part[w][threadIdx.x];
        partials[(size_t)blockIdx.x * LOSS_TERMS + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024)
loss_finalize_kernel(const float* __restrict__ partials, int nblocks, VtgsLossConfig cfg, float* __restrict__ loss_terms) {
    __shared__ double s_sum[LOSS_TERMS][256];
    __shared__ double s_tot[LOSS_TERMS];
    const int tid = threadIdx.x;
    const int term = tid & (LOSS_TERMS - 1), sl = tid >> 2;          // coalesced: consecutive threads, consecutive floats
    double acc = 0.0;
    for (int b = sl; b < nblocks; b += 256) acc += (double)partials[(size_t)b * LOSS_TERMS + term];
    s_sum[term][sl] = acc;
    __syncthreads();
    if (tid < LOSS_TERMS) {                                          // fixed order: deterministic
        double a = 0.0;
        for (int i = 0; i < 256; ++i) a += s_sum[tid][i];
        s_tot[tid] = a;
    }
    __syncthreads();
    if (tid == 0) {
        double tot[LOSS_TERMS];
        for (int k = 0; k < LOSS_TERMS; ++k) tot[k] = s_tot[k];
        const double wim = (double)cfg.w_im * tot[1], wd = (double)cfg.w_depth * tot[0];
        loss_terms[0] = (float)(wim + wd);
        loss_terms[1] = (float)wim;
        loss_terms[2] = (float)wd;
        loss_terms[3] = (float)tot[2];
        loss_terms[4] = (float)tot[1];
        loss_terms[5] = 0.0f; loss_terms[6] = 0.0f; loss_terms[7] = 0.0f;
    }
}

static inline size_t band_pixels(const CamConst& cam) {
    const size_t P = (size_t)cam.W * cam.H;
    const size_t b = (size_t)cam.row0 * 16 * cam.W;
    size_t e = (size_t)cam.row1 * 16 * cam.W;
    if (e > P) e = P;
    return e > b ? e - b : 0;
}

int launch_loss(const VtgsCamera* camera, const VtgsLossConfig* cfg, const float* image6,
                const float* gt_rgb, const float* gt_depth, float* dL_dimage4, float* loss_terms,
                float* scratch, cudaStream_t stream) {
    const CamConst cam = make_cam_const(*camera);
    if (cfg->ignore_outlier_depth) { set_error("ignore_outlier_depth_loss (median mask) is not fused"); return VTGS_E_UNSUPPORTED; }
    if (!cfg->use_l1) { set_error("use_l1 = False is not supported"); return VTGS_E_UNSUPPORTED; }
    if (cfg->mode != 0) { set_error("mapping loss (SSIM) is computed on the host side in this build"); return VTGS_E_UNSUPPORTED; }
    const size_t npx = band_pixels(cam);
    const int nblocks = (int)((npx + 255) / 256);
    if (nblocks > 0) {
        { VTGS_PROF("tracking_loss_kernel", stream); tracking_loss_kernel<<<nblocks, 256, 0, stream>>>(cam, *cfg, image6, gt_rgb, gt_depth, dL_dimage4, scratch); }
        VTGS_LAUNCH_CHECK();
    }
    { VTGS_PROF("loss_finalize_kernel", stream); loss_finalize_kernel<<<1, 1024, 0, stream>>>(scratch, nblocks, *cfg, loss_terms); }
    VTGS_LAUNCH_CHECK();
    return VTGS_OK;
}

// ---- Adam -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr, float b1, float b2, float eps, int step, const int32_t* __restrict__ step_dev) {
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        const int t = step_dev ? *step_dev : step;
        const double bc1 = 1.0 - pow((double)b1, (double)t);
        const double bc2 = 1.0 - pow((double)b2, (double)t);
        s_c[0] = (float)((double)lr / bc1);
        s_c[1] = (float)sqrt(bc2);
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = m[i] + (1.0f - b1) * (gi - m[i]);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;         // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / s_c[1] + eps;
    p[i] = p[i] - s_c[0] * (mi / denom);
}

// ---- tracking pose update: best-candidate bookkeeping + Adam on the 7 pose numbers, one launch ----
// Reference src/vtgaussian_slam.py:1890 (optimizer.step) and :1961-1970 (keep the candidate pose with
// the smallest loss; the loss of an iteration belongs to the pose BEFORE that iteration's step).
__global__ void tracking_update_kernel(float* __restrict__ cam_q, float* __restrict__ cam_t, const float* __restrict__ msg,
                                       float* __restrict__ adam, int32_t* __restrict__ step_dev, float* __restrict__ best,
                                       float lr_rot, float lr_trans, float b1, float b2, float eps) {
    const int k = threadIdx.x;           // 0..3 quaternion, 4..6 translation
    __shared__ int s_step;
    __shared__ bool s_better;
    if (k == 0) {
        s_step = *step_dev + 1;
        *step_dev = s_step;
        const float loss = msg[8];
        s_better = loss < best[0];
        if (s_better) best[0] = loss;
    }
    __syncthreads();
    if (k >= 7) return;
    float* p = k < 4 ? cam_q + k : cam_t + (k - 4);
    if (s_better) best[1 + k] = *p;
    const float g = msg[k];
    float* m = adam + (k < 4 ? k : 8 + (k - 4));
    float* v = adam + (k < 4 ? 4 + k : 11 + (k - 4));
    const double bc1 = 1.0 - pow((double)b1, (double)s_step);
    const double bc2 = 1.0 - pow((double)b2, (double)s_step);
    const float step_size = (float)((double)(k < 4 ? lr_rot : lr_trans) / bc1);
    const float mi = *m + (1.0f - b1) * (g - *m);
    const float vi = b2 * *v + (1.0f - b2) * g * g;
    *m = mi; *v = vi;
    const float denom = sqrtf(vi) / (float)sqrt(bc2) + eps;
    *p = *p - step_size * (mi / denom);
}

int launch_tracking_update(float* cam_q, float* cam_t, const float* msg, float* adam, int32_t* step_dev, float* best,
                           float lr_rot, float lr_trans, float eps, cudaStream_t stream) {
    { VTGS_PROF("tracking_update_kernel", stream); tracking_update_kernel<<<1, 32, 0, stream>>>(cam_q, cam_t, msg, adam, step_dev, best, lr_rot, lr_trans, 0.9f, 0.999f, eps); }
    VTGS_LAUNCH_CHECK();
    return VTGS_OK;
}

int launch_adam(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float b1,
                float b2, float eps, int step, const int32_t* step_dev, cudaStream_t stream) {
    if (n <= 0) return VTGS_OK;
    { VTGS_PROF("adam_kernel", stream); adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(param, grad, m, v, n, lr, b1, b2, eps, step, step_dev); }
    VTGS_LAUNCH_CHECK();
    return VTGS_OK;
}

}  // namespace vtgs

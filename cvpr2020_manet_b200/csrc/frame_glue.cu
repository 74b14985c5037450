// Frame glue after the segmentation head: what the propagation loop of the reference does between two
// prop_seghead calls.
//   test.py:253-256     pred = interpolate(pred [1,N,h,w], size=(Hf,Wf), mode='bilinear', align_corners=True); argmax(dim=1)
//   IntVOS.py:598-599   next frame: interpolate(prev_label.float(), size=(h,w), mode='nearest').int()
// One kernel evaluates the bilinear sample of all N logit planes at a full-resolution pixel and keeps the first maximum
// (torch.argmax returns the first maximal index); the [Hf,Wf] label map and the nearest-downscaled [h,w] map the next
// frame's matching consumes are both written from it -- the N x Hf x Wf upsampled logits (9.8 MB at 480p, N = 6) never exist.
// Index arithmetic follows ATen: bilinear source = dst * (in-1)/(out-1) in fp32 (area_pixel_compute_scale, align_corners),
// weights 1-lambda / lambda, horizontal blend first; nearest source = min(int(floorf(dst * (float)in/out)), in-1).
#include "common.cuh"

namespace manet {

__device__ __forceinline__ int glue_argmax_at(const float* __restrict__ logits, int N, int h, int w, float ry, float rx, int Y, int X) {
    const float sy = ry * (float)Y, sx = rx * (float)X;
    const int y1 = (int)sy, x1 = (int)sx;
    const int yp = (y1 < h - 1) ? 1 : 0, xp = (x1 < w - 1) ? 1 : 0;
    const float ly1 = sy - (float)y1, ly0 = 1.f - ly1, lx1 = sx - (float)x1, lx0 = 1.f - lx1;
    const float* p = logits + (size_t)y1 * w + x1;
    int best = 0; float bv = -INFINITY;
    for (int n = 0; n < N; ++n, p += (size_t)h * w) {
        const float v = ly0 * (lx0 * __ldg(p) + lx1 * __ldg(p + xp)) + ly1 * (lx0 * __ldg(p + (size_t)yp * w) + lx1 * __ldg(p + (size_t)yp * w + xp));
        if (v > bv) { bv = v; best = n; }
    }
    return best;
}

__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float* __restrict__ logits, int N, int h, int w, int Hf, int Wf, float ry, float rx,
                       int64_t* __restrict__ full, int32_t* __restrict__ small, float ny, float nx) {
    pdl_enter();
    const int64_t n_full = full ? (int64_t)Hf * Wf : 0, n_small = small ? (int64_t)h * w : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_full + n_small; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n_full) {
            const int Y = (int)(i / Wf), X = (int)(i % Wf);
            full[i] = glue_argmax_at(logits, N, h, w, ry, rx, Y, X);
        } else {
            const int64_t j = i - n_full;
            const int y = (int)(j / w), x = (int)(j % w);
            const int Y = min((int)floorf((float)y * ny), Hf - 1), X = min((int)floorf((float)x * nx), Wf - 1);
            small[j] = glue_argmax_at(logits, N, h, w, ry, rx, Y, X);
        }
    }
}

int launch_upsample_argmax(const float* logits, int N, int h, int w, int Hf, int Wf, int64_t* full, int32_t* small, cudaStream_t stream) {
    if (N < 1 || h < 1 || w < 1 || Hf < 1 || Wf < 1) return fail_invalid("upsample_argmax: bad sizes");
    if (!full && !small) return 0;
    const float ry = Hf > 1 ? (float)(h - 1) / (float)(Hf - 1) : 0.f, rx = Wf > 1 ? (float)(w - 1) / (float)(Wf - 1) : 0.f;
    const float ny = (float)Hf / (float)h, nx = (float)Wf / (float)w;
    const int64_t total = (full ? (int64_t)Hf * Wf : 0) + (small ? (int64_t)h * w : 0);
    const unsigned grid = (unsigned)imin64(ceil_div64(total, 256), 148 * 16);
    launch_k(upsample_argmax_kernel, dim3(grid), dim3(256), 0, stream, logits, N, h, w, Hf, Wf, ry, rx, full, small, ny, nx);
    return check_launch("upsample_argmax_kernel");
}

}  // namespace manet

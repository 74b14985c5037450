// Frame glue after the segmentation head: what the propagation loop of the reference does between two
// prop_seghead calls.
//   test.py:253-256     pred = interpolate(pred [1,N,h,w], size=(Hf,Wf), mode='bilinear', align_corners=True); argmax(dim=1)
//   IntVOS.py:598-599   next frame: interpolate(prev_label.float(), size=(h,w), mode='nearest').int()
// One kernel evaluates the bilinear sample of all N logit planes at a full-resolution pixel and keeps the first maximum
// (torch.argmax returns the first maximal index); the [Hf,Wf] label map and the nearest-downscaled [h,w] map the next
// frame's matching consumes are both written from it -- the N x Hf x Wf upsampled logits (9.8 MB at 480p, N = 6) never exist.
// Index arithmetic follows ATen: bilinear source = dst * (in-1)/(out-1) in fp32 (area_pixel_compute_scale, align_corners),
// weights 1-lambda / lambda, horizontal blend first; nearest source = min(int(floorf(dst * (float)in/out)), in-1).
#include <limits.h>

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

// ---------------------------------------------------------------------------------------------- rough_ROI
// test.py:323-343: first-round scribbles are cut to the bounding box of the scribbled pixels (label != -1) grown by
// `dist` = 20 -- inside the box labels are kept (unlabelled stays -1), outside they become 0 (background).  The reference
// finds the box with nonzero()/min()/max() and Python slicing, i.e. device->host syncs; here the box is reduced on the
// device (one atomicMin/Max set per image) and applied by a second kernel.  The slice ends are the reference's:
// rows [max(h_min-dist,0), min(h_max+dist,h-1)) and columns likewise (end exclusive, so the last row/column is never kept).
namespace manet {

__global__ void __launch_bounds__(256)
roi_bbox_kernel(const int32_t* __restrict__ labels, int B, int H, int W, int* __restrict__ box) {      // box[b] = {hmin, wmin, hmax, wmax}
    pdl_enter();
    const int b = blockIdx.y;
    const int32_t* lab = labels + (size_t)b * H * W;
    int hmin = INT_MAX, wmin = INT_MAX, hmax = -1, wmax = -1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        if (__ldg(lab + i) != -1) {
            const int y = i / W, x = i % W;
            hmin = min(hmin, y); hmax = max(hmax, y); wmin = min(wmin, x); wmax = max(wmax, x);
        }
    }
    hmin = __reduce_min_sync(0xffffffffu, hmin); wmin = __reduce_min_sync(0xffffffffu, wmin);
    hmax = __reduce_max_sync(0xffffffffu, hmax); wmax = __reduce_max_sync(0xffffffffu, wmax);
    if ((threadIdx.x & 31) == 0 && hmax >= 0) {
        atomicMin(box + 4 * b + 0, hmin); atomicMin(box + 4 * b + 1, wmin);
        atomicMax(box + 4 * b + 2, hmax); atomicMax(box + 4 * b + 3, wmax);
    }
}

__global__ void __launch_bounds__(256)
roi_apply_kernel(const int32_t* __restrict__ labels, int B, int H, int W, int dist, const int* __restrict__ box, int32_t* __restrict__ out) {
    pdl_enter();
    const int b = blockIdx.y;
    const int hmin = box[4 * b], wmin = box[4 * b + 1], hmax = box[4 * b + 2], wmax = box[4 * b + 3];
    const int y0 = max(hmin - dist, 0), y1 = min(hmax + dist, H - 1), x0 = max(wmin - dist, 0), x1 = min(wmax + dist, W - 1);
    const bool any = hmax >= 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i % W;
        const bool inside = any && y >= y0 && y < y1 && x >= x0 && x < x1;
        out[(size_t)b * H * W + i] = inside ? __ldg(labels + (size_t)b * H * W + i) : 0;
    }
}

__global__ void roi_init_kernel(int* __restrict__ box, int B) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { box[4 * i] = INT_MAX; box[4 * i + 1] = INT_MAX; box[4 * i + 2] = -1; box[4 * i + 3] = -1; }
}

int launch_rough_roi(const int32_t* labels, int B, int H, int W, int dist, int32_t* out, int* box, cudaStream_t stream) {
    if (B < 1 || H < 1 || W < 1 || dist < 0) return fail_invalid("rough_roi: bad sizes");
    launch_k(roi_init_kernel, dim3((B + 127) / 128), dim3(128), 0, stream, box, B);
    const unsigned gx = (unsigned)imin64(ceil_div64((int64_t)H * W, 256), 296);
    launch_k(roi_bbox_kernel, dim3(gx, B), dim3(256), 0, stream, labels, B, H, W, box);
    launch_k(roi_apply_kernel, dim3(gx, B), dim3(256), 0, stream, labels, B, H, W, dist, (const int*)box, out);
    return check_launch("rough_roi kernels");
}

}  // namespace manet

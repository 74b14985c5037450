"""Naive-loop second opinion for the oracle (numpy, float64 accumulation).
TEST INFRASTRUCTURE ONLY -- small shapes, pure loops, no torch.

Each function follows the *definition* of the quantity the reference computes
(SURVEY.md section 3.3-3.5), not the reference's tensor program, so that an error
shared by the reference-shaped restatement in ``manet_oracle.py`` cannot hide.

  global matching   networks/IntVOS.py:23-97, 160-210
  local matching    networks/IntVOS.py:279-296, 398-408, 428-432
  Correlation       correlation_package/correlation_cuda_kernel.cu:80-146
"""
from __future__ import annotations

import numpy as np

BIG = np.float32(1e20)


def sigmoid_norm(d):
    """(sigmoid(d) - 0.5) * 2 evaluated in float64; +inf -> 1."""
    d = np.asarray(d, dtype=np.float64)
    with np.errstate(over="ignore"):
        return (1.0 / (1.0 + np.exp(-d)) - 0.5) * 2.0


def global_match(ref, query, labels, n_ids, k=1, drop_unlabelled=False):
    """ref [R,C], query [M,C], labels [R] int; ids 0..n_ids-1.  Returns [M,n_ids]
    raw squared distances, 1e20 for an absent object."""
    ref = np.asarray(ref, np.float64)
    query = np.asarray(query, np.float64)
    labels = np.asarray(labels)
    if drop_unlabelled:
        keep = labels != -1
        ref, labels = ref[keep], labels[keep]
    out = np.empty((query.shape[0], n_ids), np.float64)
    for m in range(query.shape[0]):
        d = ((query[m][None, :] - ref) ** 2).sum(axis=1)
        for o in range(n_ids):
            cand = np.sort(d[labels == o])
            if k == 1:
                out[m, o] = cand[0] if cand.size else 1e20
            else:
                got = list(cand[:k])
                if not got:
                    # every slot is the 1e20 sentinel: mean of k copies of the filler 0*...
                    # the reference yields max(masked)=0 as filler -> 0.0
                    out[m, o] = 0.0
                else:
                    got += [got[-1]] * (k - len(got))
                    out[m, o] = float(np.mean(got))
    return out


def avg_pool2(x):
    """x [H,W,C] -> [H//2, W//2, C], 2x2 mean with floor on odd sizes."""
    h, w = x.shape[0] // 2, x.shape[1] // 2
    x = np.asarray(x, np.float64)[: 2 * h, : 2 * w]
    return (x[0::2, 0::2] + x[0::2, 1::2] + x[1::2, 0::2] + x[1::2, 1::2]) / 4.0


def bilinear_align_corners(img, out_h, out_w):
    """img [h,w] -> [out_h,out_w], align_corners=True."""
    h, w = img.shape
    out = np.empty((out_h, out_w), np.float64)
    sy = (h - 1) / (out_h - 1) if out_h > 1 else 0.0
    sx = (w - 1) / (out_w - 1) if out_w > 1 else 0.0
    for y in range(out_h):
        fy = y * sy
        y0 = min(int(np.floor(fy)), h - 1)
        y1 = min(y0 + 1, h - 1)
        wy = fy - y0
        for x in range(out_w):
            fx = x * sx
            x0 = min(int(np.floor(fx)), w - 1)
            x1 = min(x0 + 1, w - 1)
            wx = fx - x0
            top = img[y0, x0] * (1 - wx) + img[y0, x1] * wx
            bot = img[y1, x0] * (1 - wx) + img[y1, x1] * wx
            out[y, x] = top * (1 - wy) + bot * wy
    return out


def local_match(prev_emb, query_emb, prev_labels, n_ids, d):
    """prev_emb/query_emb [H,W,C], prev_labels [H,W] int.  Returns [H,W,n_ids]."""
    big_h, big_w, _ = query_emb.shape
    qs, ps = avg_pool2(query_emb), avg_pool2(prev_emb)
    h, w, _ = qs.shape
    win = 2 * d + 1
    t = np.ones((win * win, h, w), np.float64)
    for dy in range(-d, d + 1):
        for dx in range(-d, d + 1):
            l = (dy + d) * win + (dx + d)
            for y in range(h):
                yy = y + dy
                if yy < 0 or yy >= h:
                    continue
                for x in range(w):
                    xx = x + dx
                    if xx < 0 or xx >= w:
                        continue
                    diff = qs[y, x] - ps[yy, xx]
                    t[l, y, x] = sigmoid_norm(np.dot(diff, diff))
    out = np.ones((big_h, big_w, n_ids), np.float64)
    for l in range(win * win):
        dy, dx = l // win - d, l % win - d
        up = bilinear_align_corners(t[l], big_h, big_w)
        for y in range(big_h):
            yy = y + 2 * dy
            for x in range(big_w):
                xx = x + 2 * dx
                lab = prev_labels[yy, xx] if (0 <= yy < big_h and 0 <= xx < big_w) else 0
                if 0 <= lab < n_ids:
                    out[y, x, lab] = min(out[y, x, lab], up[y, x])
    return out


def correlation_forward(in1, in2, pad_size, kernel_size, max_displacement, stride1, stride2):
    """in1,in2 [B,C,H,W] -> [B,(2*dr+1)^2,outH,outW]; zero outside the padded image."""
    in1 = np.asarray(in1, np.float64)
    in2 = np.asarray(in2, np.float64)
    b, c, h, w = in1.shape
    kr = (kernel_size - 1) // 2
    dr = max_displacement // stride2
    ph, pw = h + 2 * pad_size, w + 2 * pad_size
    out_h = int(np.ceil((ph - 2 * (kr + max_displacement)) / stride1))
    out_w = int(np.ceil((pw - 2 * (kr + max_displacement)) / stride1))
    out = np.zeros((b, (2 * dr + 1) ** 2, out_h, out_w), np.float64)

    def at(img, n, yy, xx):  # padded coordinates -> value vector over channels
        y0, x0 = yy - pad_size, xx - pad_size
        if 0 <= y0 < h and 0 <= x0 < w:
            return img[n, :, y0, x0]
        return np.zeros(c)

    for n in range(b):
        for oy in range(out_h):
            for ox in range(out_w):
                y1, x1 = oy * stride1 + max_displacement, ox * stride1 + max_displacement
                for tj in range(-dr, dr + 1):
                    for ti in range(-dr, dr + 1):
                        acc = 0.0
                        for j in range(-kr, kr + 1):
                            for i in range(-kr, kr + 1):
                                acc += float(np.dot(at(in1, n, y1 + j, x1 + i),
                                                    at(in2, n, y1 + tj * stride2 + j, x1 + ti * stride2 + i)))
                        out[n, (tj + dr) * (2 * dr + 1) + ti + dr, oy, ox] = acc / (kernel_size * kernel_size * c)
    return out

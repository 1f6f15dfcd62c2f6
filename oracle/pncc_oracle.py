"""CPU restatement (numpy fp32) of the reference's PNCC rendering and refined head boxes.  TEST INFRASTRUCTURE ONLY.

  * `rasterize`      Sim3DR `_rasterize` + `get_point_weight` (head_detector/Sim3DR/lib/rasterize_kernel.cpp:54-83,219-293):
                     z-buffer over triangles in order, strict `p_depth > buffer`, barycentric weights by the
                     blackpawn dot-product formula, colour = (unsigned char)(255 * interpolated), alpha = 1.
  * `PNCCOracle`     PNCCProcessor (head_detector/pncc_processor.py:42-73): NCC colour codes of the template, faces restricted
                     to the `head_w_ears` subset, heads painted one after the other (z flipped, fresh depth buffer each).
  * `refined_head_bbox`  head_detector/utils.py:26-35.
Pinned by oracle/_ref/libsim3dr_ref.so (the reference's C++ compiled in place) in tests/test_oracle_pncc.py and by
tests/golden/pncc_ref.npz (generated with it)."""
import numpy as np

f32 = np.float32


def _weights(px, py, p0, p1, p2):
    """get_point_weight for arrays of pixel coordinates; every operation rounded to fp32 separately."""
    v0x, v0y = f32(p2[0] - p0[0]), f32(p2[1] - p0[1])
    v1x, v1y = f32(p1[0] - p0[0]), f32(p1[1] - p0[1])
    v2x, v2y = (px - p0[0]).astype(f32), (py - p0[1]).astype(f32)
    dot00 = f32(f32(v0x * v0x) + f32(v0y * v0y))
    dot01 = f32(f32(v0x * v1x) + f32(v0y * v1y))
    dot02 = (f32(v0x) * v2x).astype(f32) + (f32(v0y) * v2y).astype(f32)
    dot11 = f32(f32(v1x * v1x) + f32(v1y * v1y))
    dot12 = (f32(v1x) * v2x).astype(f32) + (f32(v1y) * v2y).astype(f32)
    den = f32(f32(dot00 * dot11) - f32(dot01 * dot01))
    inv = f32(0) if den == 0 else f32(f32(1) / den)
    u = ((dot11 * dot02).astype(f32) - (dot01 * dot12).astype(f32)).astype(f32) * inv
    v = ((dot00 * dot12).astype(f32) - (dot01 * dot02).astype(f32)).astype(f32) * inv
    u, v = u.astype(f32), v.astype(f32)
    return ((f32(1) - u).astype(f32) - v).astype(f32), v, u


def rasterize(vertices, triangles, colors, image):
    h, w, c = image.shape
    depth_buf = np.full((h, w), -1e8, dtype=f32)
    vertices = np.asarray(vertices, dtype=f32)
    colors = np.asarray(colors, dtype=f32)
    for tri in np.asarray(triangles):
        p = vertices[tri]
        x_min = max(int(np.ceil(p[:, 0].min())), 0)
        x_max = min(int(np.floor(p[:, 0].max())), w - 1)
        y_min = max(int(np.ceil(p[:, 1].min())), 0)
        y_max = min(int(np.floor(p[:, 1].max())), h - 1)
        if x_max < x_min or y_max < y_min:
            continue
        ys, xs = np.meshgrid(np.arange(y_min, y_max + 1), np.arange(x_min, x_max + 1), indexing="ij")
        w0, w1, w2 = _weights(xs.astype(f32), ys.astype(f32), p[0], p[1], p[2])
        inside = (w2 > 0) & (w1 > 0) & (w0 > 0)
        d = (((w0 * p[0, 2]).astype(f32) + (w1 * p[1, 2]).astype(f32)).astype(f32) + (w2 * p[2, 2]).astype(f32)).astype(f32)
        win = inside & (d > depth_buf[y_min:y_max + 1, x_min:x_max + 1])
        if not win.any():
            continue
        col = colors[tri]                                      # [3 vertices, c]
        pc = (((w0[..., None] * col[0]).astype(f32) + (w1[..., None] * col[1]).astype(f32)).astype(f32) + (w2[..., None] * col[2]).astype(f32)).astype(f32)
        px = (f32(255) * pc).astype(f32).astype(np.int32).astype(np.uint8)
        sub = image[y_min:y_max + 1, x_min:x_max + 1]
        sub[win] = px[win]
        depth_buf[y_min:y_max + 1, x_min:x_max + 1][win] = d[win]
    return image


def ncc_colors(v_template, subset):
    """compute_ncc_color_codes (pncc_processor.py:42-56) in the template's own dtype (float64 in the reference's asset),
    cast to fp32 where Sim3DR.rasterize does (Sim3DR.py:34-35)."""
    v_template = np.asarray(v_template)
    sub = v_template[subset]
    lo = sub.min(axis=0, keepdims=True, initial=0)
    hi = sub.max(axis=0, keepdims=True, initial=0)
    return ((v_template - lo) / (hi - lo)).astype(f32)


class PNCCOracle:
    def __init__(self, v_template, faces, head_w_ears):
        keep = np.isin(faces, head_w_ears).all(axis=1)
        self.triangles = faces[keep].astype(np.int32)
        self.colors = ncc_colors(np.asarray(v_template), np.asarray(head_w_ears))

    def __call__(self, shape_hw3, heads_vertices, raster=rasterize):
        out = np.zeros(shape_hw3, dtype=np.uint8)
        for v in heads_vertices:
            v = np.array(v, dtype=f32, copy=True)
            v[:, 2] *= -1
            cur = raster(v, self.triangles, self.colors, out.copy())
            m = cur.sum(2) != 0
            out[m] = cur[m]
        return out


def refined_head_bbox(vertices, head_indices):
    pts = np.asarray(vertices)[np.asarray(head_indices)]
    x, y, x1, y1 = (int(v) for v in (pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(), pts[:, 1].max()))
    return x, y, x1 - x, y1 - y

"""Oracle: slider end-point recompute of the diffusion `denoised_fn` (diffusion_pipeline.py:203-222).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, in plain numpy float64:
  * `SliderPath` (osuT5/osuT5/inference/slider_path.py:26-230) as the pipeline uses it — `SliderPath(curve_type, control_points)`
    with NO expected distance, then `get_distance()` and `position_at(length / max_length)`:
      - control points split into sub-paths where two consecutive points coincide (red anchors, :117-141),
      - sub-path flattening by curve type (:99-115): PerfectCurve -> circular arc only for exactly three points, else bezier;
        Catmull; everything else bezier,
      - consecutive duplicate vertices dropped (:133-139), cumulative length (:143-160), binary search + linear interpolation
        (:187-214; `binary_search` :9-23 returns an exact hit or ~insertion point).
  * `path_approximator.py`: adaptive bezier flattening by de Casteljau subdivision until every second difference is below
    BEZIER_TOLERANCE (:12-81, 173-222), circular arc with CIRCULAR_ARC_TOLERANCE (:100-161), Catmull-Rom with 50 samples per
    span (:84-97, 225-253), linear (:164-170).
  * the closure itself: in-paint, `to_positions` of the CONDITIONAL half (:172-177), per slider (fully inside the chunk) replace the
    slider-end position, then write the positions back to BOTH halves (:220, broadcast).
The reference mixes float32 control points with float64 work buffers (np.empty defaults); this restatement is float64 throughout —
pinned to the reference within 1e-3 px by tests/test_oracle_vs_reference.py on real control points of the reference's toy beatmap.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np
import torch

BEZIER_TOLERANCE = 0.25
CATMULL_DETAIL = 50
CIRCULAR_ARC_TOLERANCE = 0.1

CURVE_TYPES = {"Bezier": 0, "PerfectCurve": 1, "Catmull": 2, "Linear": 3}


@dataclasses.dataclass
class Slider:
    """`DiffusionSlider` (diffusion_pipeline.py:30-35)."""
    seq_indices: np.ndarray
    end_index: int
    curve_type: Optional[str]
    length: Optional[float]


def _flat_enough(cp: np.ndarray) -> bool:
    for i in range(1, len(cp) - 1):
        p = cp[i - 1] - 2 * cp[i] + cp[i + 1]
        if p @ p > BEZIER_TOLERANCE * BEZIER_TOLERANCE * 4:
            return False
    return True


def _subdivide(cp: np.ndarray):
    count = len(cp)
    mid = cp.copy()
    left, right = np.empty_like(cp), np.empty_like(cp)
    for i in range(count):
        left[i] = mid[0]
        right[count - i - 1] = mid[count - i - 1]
        for j in range(count - i - 1):
            mid[j] = (mid[j] + mid[j + 1]) / 2
    return left, right


def _bezier_leaf(cp: np.ndarray, out: list) -> None:
    count = len(cp)
    left, right = _subdivide(cp)
    both = np.concatenate([left, right[1:]])
    out.append(cp[0].copy())
    for i in range(1, count - 1):
        k = 2 * i
        out.append(0.25 * (both[k - 1] + 2 * both[k] + both[k + 1]))


def approximate_bezier(cp: np.ndarray) -> List[np.ndarray]:
    cp = np.asarray(cp, dtype=np.float64)
    if len(cp) == 0:
        return []
    out: list = []
    stack = [cp.copy()]
    while stack:
        parent = stack.pop()
        if _flat_enough(parent):
            _bezier_leaf(parent, out)
            continue
        left, right = _subdivide(parent)
        stack.append(right)
        stack.append(left)
    out.append(cp[-1].copy())
    return out


def _catmull_point(v1, v2, v3, v4, t):
    t2, t3 = t * t, t * t * t
    return 0.5 * (2 * v2 + (-v1 + v3) * t + (2 * v1 - 5 * v2 + 4 * v3 - v4) * t2 + (-v1 + 3 * v2 - 3 * v3 + v4) * t3)


def approximate_catmull(cp: np.ndarray) -> List[np.ndarray]:
    cp = np.asarray(cp, dtype=np.float64)
    res = []
    for i in range(len(cp) - 1):
        v1 = cp[i - 1] if i > 0 else cp[i]
        v2 = cp[i]
        v3 = cp[i + 1] if i < len(cp) - 1 else v2 + v2 - v1
        v4 = cp[i + 2] if i < len(cp) - 2 else v3 + v3 - v2
        for c in range(CATMULL_DETAIL):
            res.append(_catmull_point(v1, v2, v3, v4, c / CATMULL_DETAIL))
            res.append(_catmull_point(v1, v2, v3, v4, (c + 1) / CATMULL_DETAIL))
    return res


def approximate_circular_arc(cp: np.ndarray) -> List[np.ndarray]:
    """path_approximator.py:100-161 in the arithmetic type the pipeline feeds it: the control points are float32 (`to_positions(x)...
    .numpy()`), and every operation of the reference keeps float32 (numpy scalars; Python floats are weak).  Near-collinear
    points give radii of 1e4+ px where float32 and float64 differ by pixels, so the oracle mirrors float32 operation by operation."""
    f = np.float32
    a, b, c = (np.asarray(p, dtype=np.float32) for p in cp[:3])
    a_sq, b_sq, c_sq = f((b - c) @ (b - c)), f((a - c) @ (a - c)), f((a - b) @ (a - b))
    if np.isclose(a_sq, 0) or np.isclose(b_sq, 0) or np.isclose(c_sq, 0):
        return []
    s = a_sq * (b_sq + c_sq - a_sq)
    t = b_sq * (a_sq + c_sq - b_sq)
    u = c_sq * (a_sq + b_sq - c_sq)
    total = s + t + u
    if np.isclose(total, 0):
        return []
    centre = (s * a + t * b + u * c) / total
    d_a, d_c = a - centre, c - centre
    r = f(np.sqrt(f(d_a[0] * d_a[0] + d_a[1] * d_a[1])))
    theta_start = np.arctan2(d_a[1], d_a[0])
    theta_end = np.arctan2(d_c[1], d_c[0])
    two_pi = f(2 * np.pi)
    while theta_end < theta_start:
        theta_end = f(theta_end + two_pi)
    direction = 1
    theta_range = f(theta_end - theta_start)
    ca = c - a
    ortho = np.array([ca[1], -ca[0]], dtype=np.float32)
    if f(ortho @ (b - a)) < 0:
        direction = -1
        theta_range = f(two_pi - theta_range)
    if f(2) * r <= f(CIRCULAR_ARC_TOLERANCE):
        n = 2
    else:
        n = int(max(2, np.ceil(theta_range / (f(2) * np.arccos(f(1) - f(CIRCULAR_ARC_TOLERANCE) / r)))))
    out = []
    for i in range(n):
        theta = f(theta_start + f(direction * (i / (n - 1))) * theta_range)
        out.append((centre + np.array([np.cos(theta), np.sin(theta)], dtype=np.float32) * r).astype(np.float64))
    return out


def calculated_path(curve_type: Optional[str], control_points: np.ndarray) -> np.ndarray:
    """`SliderPath.calculate_path` (:117-141) -> (n_vertices, 2) float64."""
    cps = np.asarray(control_points, dtype=np.float64)
    n = len(cps)
    path: list = []
    start = 0
    for i in range(n):
        if i == n - 1 or (cps[i] == cps[i + 1]).all():
            span = cps[start:i + 1]
            if curve_type == "Linear":
                sub = [p.copy() for p in span]
            elif curve_type == "PerfectCurve":
                sub = approximate_circular_arc(span) if (n == 3 and len(span) == 3) else []
                if len(sub) == 0:
                    sub = approximate_bezier(span)
            elif curve_type == "Catmull":
                sub = approximate_catmull(span)
            else:
                sub = approximate_bezier(span)
            for t in sub:
                if len(path) == 0 or (path[-1] != t).any():
                    path.append(t)
            start = i + 1
    return np.array(path, dtype=np.float64).reshape(-1, 2)


def slider_end_position(curve_type: Optional[str], control_points: np.ndarray, length: float):
    """(max_length, end_pos): `SliderPath(curve_type, cps).get_distance()` and `.position_at(length / max_length)`;
    end_pos is None when max_length == 0 (the pipeline skips such sliders, diffusion_pipeline.py:215-216)."""
    path = calculated_path(curve_type, control_points)
    if len(path) == 0:
        return 0.0, None
    seg = np.linalg.norm(np.diff(path, axis=0), axis=1) if len(path) > 1 else np.zeros(0)
    cum = np.concatenate([[0.0], np.cumsum(seg)])
    max_length = float(cum[-1])
    if max_length == 0:
        return 0.0, None
    d = float(np.clip(length / max_length, 0, 1)) * max_length
    # binary_search (:9-23): exact hit -> that index, else the insertion point
    hit = np.nonzero(cum == d)[0]
    i = int(hit[0]) if len(hit) else int(np.searchsorted(cum, d, side="right"))
    if i <= 0:
        return max_length, path[0]
    if i >= len(path):
        return max_length, path[-1]
    d0, d1 = cum[i - 1], cum[i]
    if np.isclose(d0, d1):
        return max_length, path[i - 1]
    return max_length, path[i - 1] + (path[i] - path[i - 1]) * ((d - d0) / (d1 - d0))


def to_positions(x: torch.Tensor) -> np.ndarray:
    """diffusion_pipeline.py:172-177 on the conditional half: (2, 2, T) normalised -> (T, 2) float32 osu! pixels."""
    half = x[:1].clone().float()
    half += 1
    half /= 2
    half *= torch.tensor((512.0, 384.0))[None, :, None]
    return half.squeeze(0).T.numpy()


def denoised_fn_with_sliders(x: torch.Tensor, mask: torch.Tensor, z: torch.Tensor, sliders: Sequence[Slider], start: int, end: int) -> torch.Tensor:
    """The closure of `sample_part` (diffusion_pipeline.py:203-222) on CPU tensors (N = 2 rows: conditional | null class)."""
    x = torch.where(mask, x, z)
    if len(sliders) > 0:
        x2 = to_positions(x).copy()
        for s in sliders:
            if np.any((s.seq_indices < start) | (s.seq_indices >= end)) or s.end_index < start or s.end_index >= end:
                continue
            max_length, end_pos = slider_end_position(s.curve_type, x2[s.seq_indices - start], s.length)
            if max_length == 0 or end_pos is None:
                continue
            x2[s.end_index - start] = end_pos
        x = x.clone()
        x[:, :, :] = torch.from_numpy(x2.T.copy()) / torch.tensor((512, 384)).unsqueeze(1) * 2 - 1
    return x


def parse_osu_sliders(path: str):
    """Slider control points of a .osu file ([HitObjects] lines `x,y,time,type,hitSound,curveType|x:y|...,slides,length`) ->
    list of (curve_type, control_points (n, 2) float32 incl. the head, length).  Test data source only."""
    names = {"B": "Bezier", "P": "PerfectCurve", "C": "Catmull", "L": "Linear"}
    out, on = [], False
    for line in open(path, encoding="utf-8"):
        line = line.strip()
        if line.startswith("["):
            on = line == "[HitObjects]"
            continue
        if not on or not line:
            continue
        f = line.split(",")
        if len(f) < 8 or not (int(f[3]) & 2):
            continue
        parts = f[5].split("|")
        pts = [[float(f[0]), float(f[1])]] + [[float(a) for a in p.split(":")] for p in parts[1:]]
        out.append((names[parts[0]], np.array(pts, dtype=np.float32), float(f[7])))
    return out

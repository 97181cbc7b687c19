"""Second, independent float64 oracle for the nerfstudio half of the path -- TEST INFRASTRUCTURE ONLY.

oracle/oracle.py restates NeRFEncoding, the biased / uniform samplers, RaySamples.get_weights, PDFSampler and the median-depth
renderer in torch fp32, following the upstream code line by line.  A transcription error there (encoding order, searchsorted
side, include_original merge, ...) would be invisible to tests that compare the CUDA path with that restatement.  This module
computes the same quantities from their DEFINITIONS, in float64, with plain per-ray loops and different formulations
(products instead of cumulative sums, cos instead of a phase-shifted sin, linear scans instead of searchsorted, Python's
sorted() instead of a merge), never importing oracle.py.  tests/test_second_opinion.py holds the two against each other.

Definitions used (nerfstudio 0.3.x, the version the reference pins in its Dockerfile:10; call sites
tetranerf/nerfstudio/model.py:426-432, 461-463, 582-584, 632-637):
  * NeRFEncoding(in_dim=D, num_frequencies=F, min_freq_exp=a, max_freq_exp=b, include_input): for input x in R^D the output
    is [ sin(2 pi x_d 2^{e_f}) for d, f ] ++ [ cos(2 pi x_d 2^{e_f}) for d, f ] ++ x, with e_f evenly spaced on [a, b].
  * get_weights: w_j = (1 - exp(-delta_j sigma_j)) * prod_{i<j} exp(-delta_i sigma_i).
  * PDFSampler(num_samples=Sf, include_original=True, histogram_padding=0.01), eval mode: the piecewise-constant density
    (w_j + 0.01) / sum over the existing spacing bins is inverted at the Sf + 1 mid-points u_i = (i + 1/2) / (Sf + 1) (training:
    u_i = (i + r_i) / (Sf + 1), r_i uniform in [0, 1)); the new bin edges are merged with the existing ones and sorted.
  * median depth: the mid-point of the first sample at which the cumulative weight reaches 1/2 (the last sample if it never does).
  * biased sampler (model.py:111-122): the K visited tetrahedra of a ray share [0, 1] equally; a position u in [0, 1] lies in
    cell k = min(floor(u K), K - 1) at fraction u K - k and maps to near + sum_{i<k} len_i + len_k * fraction.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np


def nerf_encoding(x: Sequence[float], num_frequencies: int = 4, min_exp: float = 0.0, max_exp: float = 4.0, include_input: bool = True) -> List[float]:
    exps = [min_exp + (max_exp - min_exp) * f / (num_frequencies - 1) for f in range(num_frequencies)] if num_frequencies > 1 else [min_exp]
    sines, cosines = [], []
    for xd in x:
        for e in exps:
            arg = 2.0 * math.pi * float(xd) * (2.0 ** e)
            sines.append(math.sin(arg))
            cosines.append(math.cos(arg))
    out = sines + cosines
    if include_input:
        out += [float(v) for v in x]
    return out


def weights_from_density(deltas: Sequence[float], sigmas: Sequence[float]) -> List[float]:
    out, transmittance = [], 1.0
    for dl, sg in zip(deltas, sigmas):
        a = math.exp(-float(dl) * float(sg))
        out.append((1.0 - a) * transmittance)
        transmittance *= a
    return out


def biased_position(u: float, near: float, segments: Sequence[Sequence[float]]) -> float:
    """segments: the ray's visited tetrahedra as (t_in, t_out).  Equal share of [0,1] per tetrahedron."""
    K = len(segments)
    k = min(int(math.floor(u * K)), K - 1)
    k = max(k, 0)
    frac = u * K - k
    acc = float(near)
    for i in range(k):
        acc += max(float(segments[i][1]) - float(segments[i][0]), 0.0)
    return acc + max(float(segments[k][1]) - float(segments[k][0]), 0.0) * frac


def coarse_bins(num_samples: int, near: float, far: float, segments=None, jitter=None):
    """-> (euclidean bin edges [S+1], spacing bin edges [S+1]); segments given = biased sampler; jitter [S+1] in [0,1) =
    the stratified training draw (each edge moves inside the interval between the mid-points of its neighbouring bins)"""
    S = num_samples
    base = [j / S for j in range(S + 1)]
    if jitter is not None:
        mids = [(base[j] + base[j + 1]) / 2 for j in range(S)]
        lower, upper = [base[0]] + mids, mids + [base[-1]]
        base = [lo + (hi - lo) * float(r) for lo, hi, r in zip(lower, upper, jitter)]
    eu = [b * far + (1 - b) * near for b in base]
    if segments is None:
        return eu, base
    eu = [biased_position((e - near) / (far - near), near, segments) for e in eu]
    return eu, [(e - near) / (far - near) for e in eu]


def pdf_bins(spacing_bins: Sequence[float], weights: Sequence[float], num_fine: int, padding: float = 0.01, u_rand=None) -> List[float]:
    """-> merged, sorted spacing bin edges (len(spacing_bins) + num_fine + 1)"""
    S = len(weights)
    nb = num_fine + 1
    mass = [float(w) + padding for w in weights]
    total = sum(mass)
    if total < 1e-5:  # "eps" guard of the upstream code
        extra = 1e-5 - total
        mass = [m + extra / S for m in mass]
        total = 1e-5
    cdf = [0.0]
    for m in mass:
        cdf.append(min(1.0, cdf[-1] + m / total))
    new = []
    for i in range(nb):
        u = (i + (0.5 if u_rand is None else float(u_rand[i]))) / nb
        j = 0  # bin whose CDF interval [cdf[j], cdf[j+1]) holds u: linear scan
        while j + 1 < S and cdf[j + 1] <= u:
            j += 1
        if cdf[j + 1] <= u:  # beyond the last edge (cdf clipped at 1): sits on the last edge
            new.append(float(spacing_bins[S]))
            continue
        lo, hi = cdf[j], cdf[j + 1]
        t = (u - lo) / (hi - lo) if hi > lo else 0.0
        t = min(max(t, 0.0), 1.0)
        new.append(float(spacing_bins[j]) + t * (float(spacing_bins[j + 1]) - float(spacing_bins[j])))
    return sorted([float(b) for b in spacing_bins] + new)


def median_depth(euclid_bins: Sequence[float], weights: Sequence[float]) -> float:
    c = 0.0
    for j, w in enumerate(weights):
        c += float(w)
        if c >= 0.5:
            return (float(euclid_bins[j]) + float(euclid_bins[j + 1])) / 2
    j = len(weights) - 1
    return (float(euclid_bins[j]) + float(euclid_bins[j + 1])) / 2


def composite(euclid_bins, sigmas, colors, background=(1.0, 1.0, 1.0), clamp=True):
    """-> (rgb[3], accumulation, median depth) of one ray"""
    deltas = [float(euclid_bins[j + 1]) - float(euclid_bins[j]) for j in range(len(sigmas))]
    w = weights_from_density(deltas, sigmas)
    acc = sum(w)
    rgb = [sum(wj * float(c[ch]) for wj, c in zip(w, colors)) + background[ch] * (1.0 - acc) for ch in range(3)]
    if clamp:
        rgb = [min(max(v, 0.0), 1.0) for v in rgb]
    return rgb, acc, median_depth(euclid_bins, w)


def mlp_forward(params, features: np.ndarray, enc_dir: Sequence[float]):
    """float64 forward of the field MLP for the samples of ONE ray: features [S,64] -> (sigma [S], rgb [S,3])
    (model.py:433-455, 602-621: base 3 x (Linear + ReLU); density = softplus(Linear); colour = sigmoid(Linear(ReLU(Linear([dir, base])))))"""
    P = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    h = np.asarray(features, dtype=np.float64)
    for i in range(3):
        h = np.maximum(h @ P[f"mlp_base.layers.{i}.weight"].T + P[f"mlp_base.layers.{i}.bias"], 0.0)
    s = h @ P["field_output_density.net.weight"].T + P["field_output_density.net.bias"]
    sigma = np.where(s > 20.0, s, np.log1p(np.exp(np.minimum(s, 20.0))))[:, 0]
    x = np.concatenate([np.broadcast_to(np.asarray(enc_dir, dtype=np.float64), (h.shape[0], len(enc_dir))), h], axis=1)
    g = np.maximum(x @ P["mlp_head.layers.0.weight"].T + P["mlp_head.layers.0.bias"], 0.0)
    z = g @ P["field_output_color.net.weight"].T + P["field_output_color.net.bias"]
    return sigma, 1.0 / (1.0 + np.exp(-z))

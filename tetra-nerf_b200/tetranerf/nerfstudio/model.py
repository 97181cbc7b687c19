"""Host-side mirror of the reference's nerfstudio plug-in for the ray-sampling hot path.

Same public names, constructor arguments, config fields, state-dict keys/shapes and output dictionary as
the reference `tetranerf/nerfstudio/model.py` (TetrahedraNerfConfig :70-107, TetrahedraSampler :125-192,
TetrahedraNerf :209-662), so `ns-train tetra-nerf` can import this module unchanged.  What differs is where
the work happens:

  * inference (`not self.training` and no autograd): the whole body of `get_outputs` after the ray bundle --
    trace, sampling, matching, interpolation, both MLP passes, PDF resampling and compositing
    (reference :526-662) -- is ONE call into the fused CUDA pipeline (`tetranerf.b200.render.FusedRenderer`);
  * training: the reference's own sequence of calls (trace_rays -> sampler -> find_visited_cells ->
    interpolate_values -> torch MLPs -> renderers) is kept, on top of our CUDA ops, so that autograd
    reaches `tetrahedra_field` and the MLP weights exactly as it does upstream.

Evaluation images / metrics (`get_image_metrics_and_images`, reference :676-713) are provided with PSNR and SSIM computed in torch
(torchmetrics / LPIPS are used when nerfstudio and its dependencies are installed); appearance embeddings (reference :440-446,
608-620) run on the unfused path.  Out of scope and absent: the unused occupancy field (:98,256-265).  nerfstudio itself is imported when it is
installed; otherwise the minimal look-alikes of `_ns_compat` are used.
"""
from __future__ import annotations

import dataclasses
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Literal, Optional

import torch
from torch import nn
from torch.nn import Parameter

try:  # the real thing when available (reference model.py:10-28)
    from nerfstudio.cameras.rays import RayBundle, RaySamples
    from nerfstudio.field_components.encodings import NeRFEncoding
    from nerfstudio.field_components.field_heads import DensityFieldHead, FieldHeadNames, RGBFieldHead
    from nerfstudio.field_components.mlp import MLP
    from nerfstudio.model_components.losses import MSELoss
    from nerfstudio.model_components.ray_samplers import PDFSampler, Sampler, UniformSampler
    from nerfstudio.model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer
    from nerfstudio.models.base_model import Model, ModelConfig
    from nerfstudio.utils.misc import scale_dict

    HAVE_NERFSTUDIO = True
except ImportError:
    from ._ns_compat import (AccumulationRenderer, DensityFieldHead, DepthRenderer, FieldHeadNames, MLP, MSELoss, Model, ModelConfig,
                             NeRFEncoding, PDFSampler, RayBundle, RaySamples, RGBFieldHead, RGBRenderer, Sampler, UniformSampler, scale_dict)

    HAVE_NERFSTUDIO = False

from ..utils.extension import TetrahedraTracer, interpolate_values, triangulate


@dataclass
class TetrahedraNerfConfig(ModelConfig):
    """Field-for-field the reference config (model.py:70-107)."""

    _target: Any = dataclasses.field(default_factory=lambda: TetrahedraNerf)
    tetrahedra_path: Optional[Path] = None
    num_tetrahedra_vertices: Optional[int] = None
    num_tetrahedra_cells: Optional[int] = None
    max_intersected_triangles: int = 512
    num_samples: int = 256
    num_fine_samples: int = 256
    use_biased_sampler: bool = False
    field_dim: int = 64
    num_color_layers: int = 1
    num_density_layers: int = 3
    hidden_size: int = 128
    input_fourier_frequencies: int = 0
    initialize_colors: bool = True
    use_gradient_scaling: bool = False
    background_color: Literal["random", "last_sample", "black", "white"] = "white"
    appearance_embed_dim: int = 0
    use_occupancy_field: bool = False

    def __post_init__(self):
        if self.tetrahedra_path is not None and self.num_tetrahedra_vertices is None:
            if not Path(self.tetrahedra_path).exists():
                raise RuntimeError(f"Tetrahedra path {self.tetrahedra_path} does not exist")
            th = torch.load(self.tetrahedra_path)
            self.num_tetrahedra_vertices = len(th["vertices"])
            self.num_tetrahedra_cells = len(th["cells"])


def map_from_real_distances_to_biased_with_bounds(num_bounds, bounds, samples):
    """Every visited tetrahedron receives the same share of the unit interval (reference model.py:111-122).

    num_bounds i64[R]; bounds f32[R,M,2] = (t_in, t_out) per visited cell; samples f32[R,S] euclidean distances in
    [first t_in, last t_out].  Returns the samples re-mapped onto the concatenated cell intervals."""
    seg_len = (bounds[..., 1] - bounds[..., 0]).clamp_min(0)
    first = bounds[:, 0, 0]
    last = torch.gather(bounds[..., 1], 1, (num_bounds[:, None] - 1).clamp_min(0)).squeeze(-1)
    pos = (samples - first[:, None]) / (last - first)[:, None] * num_bounds[:, None]
    cell = torch.minimum(pos.floor(), (num_bounds[:, None] - 1).to(pos.dtype)).clamp_min(0)
    frac = pos - cell
    cell = cell.long()
    starts = torch.cumsum(torch.cat((first[:, None], seg_len), 1), 1)
    return torch.gather(starts, 1, cell) + torch.gather(seg_len, 1, cell) * frac


class TetrahedraSampler(Sampler):
    """Biased sampler: an equal number of bins per visited tetrahedron (reference model.py:125-192)."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True) -> None:
        super().__init__(num_samples=num_samples)
        self.train_stratified = train_stratified

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None, *, num_visited_cells, hit_distances) -> RaySamples:
        assert ray_bundle is not None and ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        dev = ray_bundle.origins.device
        bins = torch.linspace(0.0, 1.0, num_samples + 1).to(dev)[None, ...]
        if self.train_stratified and self.training:  # per-bin jitter, as nerfstudio's SpacedSampler
            jitter = torch.rand((ray_bundle.origins.shape[0], num_samples + 1), dtype=bins.dtype, device=dev)
            mids = (bins[..., 1:] + bins[..., :-1]) / 2.0
            hi = torch.cat([mids, bins[..., -1:]], -1)
            lo = torch.cat([bins[..., :1], mids], -1)
            bins = lo + (hi - lo) * jitter
        near, far = ray_bundle.nears, ray_bundle.fars

        def to_euclidean(x):
            return x * far + (1 - x) * near

        euclid = map_from_real_distances_to_biased_with_bounds(num_visited_cells.long(), hit_distances, to_euclidean(bins))
        bins = (euclid - near) / (far - near)
        return ray_bundle.get_ray_samples(bin_starts=euclid[..., :-1, None], bin_ends=euclid[..., 1:, None], spacing_starts=bins[..., :-1, None],
                                          spacing_ends=bins[..., 1:, None], spacing_to_euclidean_fn=to_euclidean)


class GradientScaler(torch.autograd.Function):
    """Near-camera gradient damping, squared ray distance clamped to [0,1] (reference model.py:195-205)."""

    @staticmethod
    def forward(ctx, colors, sigmas, ray_dist):
        ctx.save_for_backward(ray_dist)
        return colors, sigmas, ray_dist

    @staticmethod
    def backward(ctx, g_colors, g_sigmas, g_dist):
        (ray_dist,) = ctx.saved_tensors
        s = torch.square(ray_dist).clamp(0, 1)
        return g_colors * s, g_sigmas * s, g_dist


class TetrahedraNerf(Model):
    """Tetra-NeRF model on the B200-native tracer.  Buffers `tetrahedra_vertices` f32[V,3], `tetrahedra_cells`
    i32[T,4] and parameter `tetrahedra_field` f32[field_dim,V] keep the reference's names and shapes
    (model.py:239-255) so checkpoints interchange."""

    config: TetrahedraNerfConfig

    def __init__(self, config: TetrahedraNerfConfig, dataparser_transform=None, dataparser_scale=None, metadata=None, **kwargs) -> None:
        super().__init__(config=config, **kwargs)
        self.dataparser_transform = dataparser_transform
        self.dataparser_scale = dataparser_scale
        self._tetrahedra_tracer = None
        self._fused = None
        self._fused_versions = None
        if self.config.tetrahedra_path is None and metadata is not None and "points3D_xyz" in metadata:
            self._load_points_from_metadata(**metadata)
        else:
            if self.config.num_tetrahedra_vertices is None or self.config.num_tetrahedra_cells is None:
                raise RuntimeError("The tetrahedra_path must be specified.")
            V, T = self.config.num_tetrahedra_vertices, self.config.num_tetrahedra_cells
            self.register_buffer("tetrahedra_vertices", torch.empty((V, 3), dtype=torch.float32))
            self.register_buffer("tetrahedra_cells", torch.empty((T, 4), dtype=torch.int32))
            self.register_parameter("tetrahedra_field", nn.Parameter(torch.empty((self.config.field_dim, V), dtype=torch.float32)))
            self._tetrahedra_initialized = False

    # ---- initialisation (reference :268-392) --------------------------------------------------------
    @staticmethod
    def _init_tetrahedra_field(tetrahedra_field):
        tetrahedra_field.uniform_(-1e-4, 1e-4)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        complete = all(f"{prefix}{k}" in state_dict for k in ("tetrahedra_vertices", "tetrahedra_cells", "tetrahedra_field"))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        if complete:
            self._tetrahedra_initialized = True

    def _install_mesh(self, vertices: torch.Tensor, cells: torch.Tensor, colors: Optional[torch.Tensor], alpha: Optional[torch.Tensor]):
        V = len(vertices)
        self.config.num_tetrahedra_vertices, self.config.num_tetrahedra_cells = V, len(cells)
        if hasattr(self, "tetrahedra_vertices"):
            self.tetrahedra_vertices.copy_(vertices.to(self.tetrahedra_vertices.device))
            self.tetrahedra_cells.copy_(cells.to(torch.int32).to(self.tetrahedra_cells.device))
        else:
            self.register_buffer("tetrahedra_vertices", vertices.float())
            self.register_buffer("tetrahedra_cells", cells.to(torch.int32))
            self.register_parameter("tetrahedra_field", nn.Parameter(torch.empty((self.config.field_dim, V), dtype=torch.float32)))
        self._init_tetrahedra_field(self.tetrahedra_field.data)
        if self.config.initialize_colors:
            assert colors is not None and colors.dtype == torch.uint8
            rgb = colors.float().to(self.tetrahedra_field.device) * 2.0 / 255.0 - 1.0
            self.tetrahedra_field.data[1:4, :] = rgb[:, :3].T
            self.tetrahedra_field.data[0, :] = 1.0 if alpha is None else alpha.float().to(self.tetrahedra_field.device) * 2.0 / 255.0 - 1.0
        self._tetrahedra_initialized = True

    def _load_points_from_metadata(self, points3D_xyz, points3D_rgb=None, **kwargs):
        cells = triangulate(points3D_xyz).int()
        self._install_mesh(points3D_xyz, cells, points3D_rgb, None)

    def _init_tetrahedra(self):
        if self.config.tetrahedra_path is None:
            raise RuntimeError("The tetrahedra_path must be specified.")
        path = Path(self.config.tetrahedra_path)
        if not path.exists():
            raise RuntimeError(f"Specified tetrahedra path {path} does not exist")
        if self.dataparser_scale is None:
            raise RuntimeError("Could not read the dataparser_scale and dataparser_transform parameters."
                               "Make sure you are using the TetrahedraNerfPipeline with the model.")
        th = torch.load(str(path), map_location=torch.device("cpu"))  # {"vertices", "cells", "colors"} (scripts/triangulate.py:68-75)
        verts = th["vertices"].float()
        verts = torch.cat((verts, torch.ones_like(verts[..., :1])), -1) @ self.dataparser_transform.T
        verts = verts * self.dataparser_scale
        colors = th.get("colors")
        self._install_mesh(verts, th["cells"].int(), colors, colors[:, 3] if colors is not None else None)

    def get_tetrahedra_tracer(self):
        device = self.tetrahedra_field.device
        if device.type != "cuda":
            raise RuntimeError("Tetrahedra tracer is only supported on a CUDA device")  # reference :396-397
        if self._tetrahedra_tracer is not None and self._tetrahedra_tracer.device != device:
            self._tetrahedra_tracer = self._fused = self._fused_versions = None
        if self._tetrahedra_tracer is None:
            if not self._tetrahedra_initialized:
                self._init_tetrahedra()
            self._tetrahedra_tracer = TetrahedraTracer(device)
            self._tetrahedra_tracer.load_tetrahedra(self.tetrahedra_vertices, self.tetrahedra_cells)
        return self._tetrahedra_tracer

    # ---- modules (reference :409-477) ----------------------------------------------------------------
    def populate_modules(self):
        super().populate_modules()
        in_dim = self.config.field_dim
        if self.config.input_fourier_frequencies > 0:
            self.position_encoding = NeRFEncoding(in_dim=in_dim, num_frequencies=self.config.input_fourier_frequencies, min_freq_exp=0.0,
                                                  max_freq_exp=float(self.config.input_fourier_frequencies), include_input=True)
            in_dim += self.position_encoding.get_out_dim()
        else:
            self.position_encoding = lambda x: x
        self.direction_encoding = NeRFEncoding(in_dim=3, num_frequencies=4, min_freq_exp=0.0, max_freq_exp=4.0, include_input=True)
        self.mlp_base = MLP(in_dim=in_dim, num_layers=self.config.num_density_layers, layer_width=self.config.hidden_size, out_activation=nn.ReLU())
        head_in = self.mlp_base.get_out_dim() + self.direction_encoding.get_out_dim()
        if self.config.appearance_embed_dim > 0:  # reference :440-446
            self.appearance_embedding = nn.Embedding(self.num_train_data, self.config.appearance_embed_dim)
            head_in += self.config.appearance_embed_dim
        self.mlp_head = MLP(in_dim=head_in, num_layers=self.config.num_color_layers, layer_width=self.config.hidden_size, out_activation=nn.ReLU())
        self.field_output_color = RGBFieldHead(in_dim=self.mlp_head.get_out_dim())
        self.field_output_density = DensityFieldHead(in_dim=self.mlp_base.get_out_dim())
        if self.config.use_biased_sampler:
            self.sampler_uniform = TetrahedraSampler(num_samples=self.config.num_samples)
        else:
            self.sampler_uniform = UniformSampler(num_samples=self.config.num_samples)
        if self.config.num_fine_samples > 0:
            self.sampler_pdf = PDFSampler(num_samples=self.config.num_fine_samples)
        self.renderer_rgb = RGBRenderer(background_color=self.config.background_color)
        self.renderer_accumulation = AccumulationRenderer()
        self.renderer_depth = DepthRenderer()
        self.rgb_loss = MSELoss()

    def get_param_groups(self) -> Dict[str, List[Parameter]]:
        return {"fields": list(self.parameters())}

    def get_background_color(self, shape, device):
        return self.renderer_rgb.get_background_color(self.renderer_rgb.background_color, shape, device)

    # ---- fused inference path ---------------------------------------------------------------------------
    def _fused_supported(self) -> bool:
        c = self.config
        return (c.field_dim == 64 and c.hidden_size == 128 and c.num_density_layers == 3 and c.num_color_layers == 1
                and c.input_fourier_frequencies == 0 and c.appearance_embed_dim == 0 and c.background_color in ("white", "black"))

    def _fused_renderer(self):
        from ..b200.render import FusedRenderer

        tracer = self.get_tetrahedra_tracer()
        if self._fused is None:
            self._fused = FusedRenderer(tracer)
        names = ["mlp_base.layers.0", "mlp_base.layers.1", "mlp_base.layers.2", "mlp_head.layers.0", "field_output_color.net", "field_output_density.net"]
        mods = [self.mlp_base.layers[0], self.mlp_base.layers[1], self.mlp_base.layers[2], self.mlp_head.layers[0], self.field_output_color.net,
                self.field_output_density.net]
        versions = (self.tetrahedra_field._version, self.tetrahedra_field.data_ptr()) + tuple(p._version for m in mods for p in (m.weight, m.bias))
        if versions != self._fused_versions:  # refresh the [V,64] shadow / packed weights only when a parameter changed
            self._fused.set_field(self.tetrahedra_field.detach().contiguous())
            self._fused.set_weights({f"{n}.{k}": getattr(m, k) for n, m in zip(names, mods) for k in ("weight", "bias")})
            self._fused_versions = versions
        return self._fused

    # ---- forward (reference :520-662) ---------------------------------------------------------------------
    def get_outputs(self, ray_bundle: RayBundle):
        assert self.collider is not None
        origins, directions = ray_bundle.origins.contiguous(), ray_bundle.directions.contiguous()
        if not self.training and not torch.is_grad_enabled() and self._fused_supported():
            from ..b200.render import RenderSettings

            bg = (1.0, 1.0, 1.0) if self.config.background_color == "white" else (0.0, 0.0, 0.0)
            st = RenderSettings(self.config.max_intersected_triangles, self.config.num_samples, self.config.num_fine_samples,
                                self.config.use_biased_sampler, float(self.collider.far_plane), bg)
            return self._fused_renderer().render(origins, directions, st)
        if self.training and torch.is_grad_enabled() and self._fused_supported() and self.config.num_fine_samples > 0 \
                and os.environ.get("TETRANERF_B200_UNFUSED_TRAIN", "0") != "1":
            return self._get_outputs_fused_train(origins, directions)
        return self._get_outputs_unfused(ray_bundle, origins, directions)

    def _get_outputs_fused_train(self, origins, directions):
        """training step on the fused CUDA pipeline: ONE differentiable op (forward + tcgen05 backward) instead of the reference's op
        sequence; the stratified draws are the same two torch.rand calls the reference makes (model.py:169-174, PDFSampler)."""
        from ..b200.render import PARAM_ORDER, FusedTrainRender, RenderSettings

        c = self.config
        bg = (1.0, 1.0, 1.0) if c.background_color == "white" else (0.0, 0.0, 0.0)
        st = RenderSettings(c.max_intersected_triangles, c.num_samples, c.num_fine_samples, c.use_biased_sampler, float(self.collider.far_plane), bg)
        fr = self._fused_renderer()
        R, dev = origins.shape[0], origins.device
        jc = torch.rand((R, c.num_samples + 1), dtype=torch.float32, device=dev) if getattr(self.sampler_uniform, "train_stratified", True) else None
        jf = torch.rand((R, c.num_fine_samples + 1), dtype=torch.float32, device=dev) if getattr(self.sampler_pdf, "train_stratified", True) else None
        named = dict(self.named_parameters())
        rgb, acc, depth, mask = FusedTrainRender.apply(fr, st, c.use_gradient_scaling, origins, directions, jc, jf, self.tetrahedra_field,
                                                       *[named[n] for n in PARAM_ORDER])
        return {"rgb": rgb, "accumulation": acc, "depth": depth, "ray_mask": mask}

    def _field_at(self, tracer, traced, ray_mask, distances):
        matched = tracer.find_visited_cells(traced["num_visited_cells"][ray_mask], traced["visited_cells"][ray_mask],
                                            traced["barycentric_coordinates"][ray_mask], traced["hit_distances"][ray_mask],
                                            traced["vertex_indices"][ray_mask], distances.squeeze(-1).contiguous())
        return interpolate_values(matched["vertex_indices"], matched["barycentric_coordinates"], self.tetrahedra_field)

    def _get_outputs_unfused(self, ray_bundle, origins, directions):
        tracer = self.get_tetrahedra_tracer()
        traced = tracer.trace_rays(origins, directions, self.config.max_intersected_triangles)
        count = traced["num_visited_cells"]
        nears = traced["hit_distances"][:, 0, 0][:, None]
        fars = torch.gather(traced["hit_distances"][:, :, 1], 1, (count[:, None].long() - 1).clamp_min(0))
        ray_mask = count > 0
        device = ray_mask.device
        R = ray_mask.shape[0]
        rgb = self.get_background_color((R, 3), device=device)
        accumulation = torch.zeros((R, 1), dtype=torch.float32, device=device)
        depth = torch.full((R, 1), self.collider.far_plane, dtype=torch.float32, device=device)
        if bool(ray_mask.any()):
            bundle = dataclasses.replace(ray_bundle[ray_mask], nears=nears[ray_mask], fars=fars[ray_mask])
            if isinstance(self.sampler_uniform, TetrahedraSampler):
                samples = self.sampler_uniform(bundle, num_visited_cells=count[ray_mask], hit_distances=traced["hit_distances"][ray_mask])
            else:
                samples = self.sampler_uniform(bundle)
            features = self._field_at(tracer, traced, ray_mask, (samples.frustums.ends + samples.frustums.starts) / 2)
            if self.config.num_fine_samples > 0:
                coarse_density = self.field_output_density(self.mlp_base(self.position_encoding(features)))
                samples = self.sampler_pdf(bundle, samples, samples.get_weights(coarse_density))
                features = self._field_at(tracer, traced, ray_mask, (samples.frustums.ends + samples.frustums.starts) / 2)
            base = self.mlp_base(self.position_encoding(features))
            sigmas = self.field_output_density(base)
            head_in = [self.direction_encoding(samples.frustums.directions), base]
            if self.config.appearance_embed_dim > 0:  # reference :608-620
                if self.training:
                    assert samples.camera_indices is not None
                    head_in.append(self.appearance_embedding(samples.camera_indices.squeeze()))
                else:
                    head_in.append(torch.ones((*base.shape[:-1], self.config.appearance_embed_dim), device=base.device) * self.appearance_embedding.weight.mean(dim=0))
            colors = self.field_output_color(self.mlp_head(torch.cat(head_in, dim=-1)))
            if self.config.use_gradient_scaling:
                colors, sigmas, _ = GradientScaler.apply(colors, sigmas, samples.spacing_ends + samples.spacing_starts)
            weights = samples.get_weights(sigmas)
            rgb[ray_mask] = self.renderer_rgb(rgb=colors, weights=weights)
            accumulation[ray_mask] = self.renderer_accumulation(weights)
            depth[ray_mask] = self.renderer_depth(weights, samples)
        return {"rgb": rgb, "accumulation": accumulation, "depth": depth, "ray_mask": ray_mask}

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, torch.Tensor]:
        image = batch["image"].to(outputs["rgb"].device)
        return scale_dict({"rgb_loss": self.rgb_loss(image, outputs["rgb"])}, self.config.loss_coefficients)

    # ---- evaluation images and metrics (reference :676-713) --------------------------------------------------------------------------
    def get_image_metrics_and_images(self, outputs: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor]):
        image = batch["image"].to(outputs["rgb"].device)
        rgb = outputs["rgb"]
        acc = _apply_colormap(outputs["accumulation"])
        depth = _apply_depth_colormap(outputs["depth"], accumulation=outputs["accumulation"])
        images = {"img": torch.cat([image, rgb], dim=1), "accumulation": torch.cat([acc], dim=1), "depth": torch.cat([depth], dim=1)}
        # [H, W, C] -> [1, C, H, W]
        im, pr = torch.moveaxis(image, -1, 0)[None, ...], torch.moveaxis(rgb, -1, 0)[None, ...]
        metrics = {"psnr": float(_psnr(im, pr)), "nerfstudio_ssim": float(_ssim(im, pr))}
        lp = _lpips(im, pr)
        if lp is not None:
            metrics["lpips"] = float(lp)
        return metrics, images


def _psnr(a: torch.Tensor, b: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """torchmetrics PeakSignalNoiseRatio(data_range=1.0) (reference :470)"""
    mse = torch.mean((a - b) ** 2)
    return 10.0 * torch.log10(data_range**2 / mse)


def _ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 1.0, kernel: int = 11, sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03) -> torch.Tensor:
    """torchmetrics.functional.structural_similarity_index_measure with its defaults (Gaussian 11x11, sigma 1.5), [N,C,H,W] in [0,1]"""
    c = a.shape[1]
    x = torch.arange(kernel, dtype=a.dtype, device=a.device) - (kernel - 1) / 2
    g = torch.exp(-(x**2) / (2 * sigma**2))
    g = (g / g.sum())[:, None] * (g / g.sum())[None, :]
    w = g.expand(c, 1, kernel, kernel).contiguous()
    pad = (kernel - 1) // 2
    ap, bp = (torch.nn.functional.pad(t, (pad, pad, pad, pad), mode="reflect") for t in (a, b))

    def f(t):
        return torch.nn.functional.conv2d(t, w, groups=c)

    mu_a, mu_b = f(ap), f(bp)
    s_aa, s_bb, s_ab = f(ap * ap) - mu_a**2, f(bp * bp) - mu_b**2, f(ap * bp) - mu_a * mu_b
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    ssim = ((2 * mu_a * mu_b + c1) * (2 * s_ab + c2)) / ((mu_a**2 + mu_b**2 + c1) * (s_aa + s_bb + c2))
    return ssim.mean()


_LPIPS = None


def _lpips(a, b):
    """LearnedPerceptualImagePatchSimilarity (reference :474) when torchmetrics + its weights are available; None otherwise"""
    global _LPIPS
    if _LPIPS is False:
        return None
    try:
        if _LPIPS is None:
            from torchmetrics.image.lpip import LearnedPerceptualImagePatchSimilarity

            _LPIPS = LearnedPerceptualImagePatchSimilarity().to(a.device)
        return _LPIPS(a, b)
    except Exception:  # not installed / no pretrained weights offline
        _LPIPS = False
        return None


def _apply_colormap(x: torch.Tensor) -> torch.Tensor:
    try:
        from nerfstudio.utils import colormaps

        return colormaps.apply_colormap(x)
    except ImportError:
        return torch.nan_to_num(x, 0.0).clamp(0, 1).expand(*x.shape[:-1], 3)


def _apply_depth_colormap(depth: torch.Tensor, accumulation: Optional[torch.Tensor] = None) -> torch.Tensor:
    try:
        from nerfstudio.utils import colormaps

        return colormaps.apply_depth_colormap(depth, accumulation=accumulation)
    except ImportError:
        near, far = float(depth.min()), float(depth.max())
        d = ((depth - near) / (far - near + 1e-10)).clamp(0, 1)
        img = torch.nan_to_num(d, 0.0).expand(*d.shape[:-1], 3)
        return img * accumulation + (1 - accumulation) if accumulation is not None else img

"""CPU tests of the host-side mirror of the nerfstudio plug-in (config, sampler maths, state-dict contract)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tetranerf.b200 import synthetic as syn
from tetranerf.nerfstudio import model as M


def test_config_defaults_match_reference():
    c = M.TetrahedraNerfConfig(num_tetrahedra_vertices=10, num_tetrahedra_cells=5)  # model.py:70-107
    assert (c.max_intersected_triangles, c.num_samples, c.num_fine_samples, c.use_biased_sampler) == (512, 256, 256, False)
    assert (c.field_dim, c.num_color_layers, c.num_density_layers, c.hidden_size) == (64, 1, 3, 128)
    assert c.background_color == "white" and c.input_fourier_frequencies == 0 and not c.use_gradient_scaling


def test_state_dict_contract():
    m = M.TetrahedraNerf(M.TetrahedraNerfConfig(num_tetrahedra_vertices=7, num_tetrahedra_cells=3))
    sd = m.state_dict()
    assert sd["tetrahedra_vertices"].shape == (7, 3) and sd["tetrahedra_vertices"].dtype == torch.float32
    assert sd["tetrahedra_cells"].shape == (3, 4) and sd["tetrahedra_cells"].dtype == torch.int32
    assert sd["tetrahedra_field"].shape == (64, 7)  # feature-major (model.py:247-255)
    for k, shape in {"mlp_base.layers.0.weight": (128, 64), "mlp_base.layers.2.weight": (128, 128), "mlp_head.layers.0.weight": (128, 155),
                     "field_output_color.net.weight": (3, 128), "field_output_density.net.weight": (1, 128)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert not m._tetrahedra_initialized
    m2 = M.TetrahedraNerf(M.TetrahedraNerfConfig(num_tetrahedra_vertices=7, num_tetrahedra_cells=3))
    m2.load_state_dict(sd)
    assert m2._tetrahedra_initialized  # model.py:273-300
    with pytest.raises(RuntimeError, match="CUDA"):
        m2.get_tetrahedra_tracer()  # model.py:394-397


def test_biased_mapping_and_sampler_against_oracle(small_mesh):
    V, C = small_mesh
    mesh = orc.OracleMesh(V, C)
    o, d = syn.camera_rays(40)
    tr = mesh.trace_rays(o, d, 256)
    keep = tr["num_visited_cells"] > 0
    num = torch.from_numpy(tr["num_visited_cells"][keep])
    hd = torch.from_numpy(tr["hit_distances"][keep])
    nears = hd[:, 0, 0][:, None]
    fars = torch.gather(hd[:, :, 1], 1, (num[:, None].long() - 1).clamp_min(0))
    cfg = orc.RenderConfig.tetra_nerf()
    ref_e, ref_s = orc.coarse_bins(cfg, nears, fars, num, hd)
    sampler = M.TetrahedraSampler(num_samples=128).eval()
    bundle = M.RayBundle(origins=torch.from_numpy(o[keep]), directions=torch.from_numpy(d[keep]), nears=nears, fars=fars)
    rs = sampler(bundle, num_visited_cells=num, hit_distances=hd)
    got_e = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[:, -1:, 0]], -1)
    torch.testing.assert_close(got_e, ref_e, rtol=0, atol=0)
    torch.testing.assert_close(torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[:, -1:, 0]], -1), ref_s, rtol=0, atol=0)
    # equal number of bins per visited cell: bin edges that fall on cell boundaries
    k = int(num[0])
    edges = got_e[0]
    assert abs(float(edges[0]) - float(hd[0, 0, 0])) < 1e-6 and abs(float(edges[-1]) - float(hd[0, k - 1, 1])) < 1e-5
    # training mode jitters but stays sorted and inside [near, far]
    sampler.train()
    torch.manual_seed(0)
    rs2 = sampler(bundle, num_visited_cells=num, hit_distances=hd)
    e2 = torch.cat([rs2.frustums.starts[..., 0], rs2.frustums.ends[:, -1:, 0]], -1)
    assert (e2[:, 1:] >= e2[:, :-1] - 1e-6).all() and (e2 >= nears - 1e-5).all() and (e2 <= fars + 1e-5).all()


def test_compat_layers_match_oracle_restatement():
    if M.HAVE_NERFSTUDIO:
        pytest.skip("real nerfstudio present")
    torch.manual_seed(0)
    enc = M.NeRFEncoding(in_dim=3, num_frequencies=4, min_freq_exp=0.0, max_freq_exp=4.0, include_input=True)
    x = torch.nn.functional.normalize(torch.randn(10, 3), dim=-1)
    assert enc.get_out_dim() == 27
    torch.testing.assert_close(enc(x), orc.nerf_encoding_dirs(x))
    m = M.TetrahedraNerf(M.TetrahedraNerfConfig(num_tetrahedra_vertices=4, num_tetrahedra_cells=1))
    p = {k: v for k, v in m.state_dict().items() if k.startswith(("mlp_", "field_output"))}
    f = torch.randn(5, 9, 64)
    torch.testing.assert_close(m.mlp_base(f), orc.mlp_base(p, f))
    torch.testing.assert_close(m.field_output_density(m.mlp_base(f)), orc.density_head(p, orc.mlp_base(p, f)))


def test_th_file_loader_and_dataparser_transform(tmp_path, small_mesh):
    """`.th` input ({"vertices","cells","colors"}, scripts/triangulate.py:68-75) -> config sizes (model.py:102-107),
    vertices through the dataparser transform + scale, colour/alpha initialisation of the field (model.py:343-392)."""
    V, C = small_mesh
    rng = np.random.default_rng(3)
    colors = rng.integers(0, 256, size=(len(V), 4), dtype=np.uint8)
    path = tmp_path / "scene.th"
    torch.save({"vertices": torch.from_numpy(V), "cells": torch.from_numpy(C.astype(np.int32)), "colors": torch.from_numpy(colors)}, path)
    cfg = M.TetrahedraNerfConfig(tetrahedra_path=path)
    assert (cfg.num_tetrahedra_vertices, cfg.num_tetrahedra_cells) == (len(V), len(C))
    with pytest.raises(RuntimeError, match="does not exist"):
        M.TetrahedraNerfConfig(tetrahedra_path=tmp_path / "missing.th")
    T = torch.tensor([[0.0, -1.0, 0.0, 0.5], [1.0, 0.0, 0.0, -0.25], [0.0, 0.0, 1.0, 2.0]])  # [3,4] rotation + translation
    m = M.TetrahedraNerf(cfg, dataparser_transform=T, dataparser_scale=0.5)
    assert not m._tetrahedra_initialized
    m._init_tetrahedra()
    want = (np.concatenate([V, np.ones((len(V), 1), np.float32)], 1) @ T.numpy().T) * 0.5
    assert np.allclose(m.tetrahedra_vertices.numpy(), want, atol=1e-6)
    assert np.array_equal(m.tetrahedra_cells.numpy(), C.astype(np.int32))
    f = m.tetrahedra_field.detach().numpy()
    assert np.allclose(f[1:4].T, colors[:, :3].astype(np.float32) * 2 / 255 - 1, atol=1e-6)
    assert np.allclose(f[0], colors[:, 3].astype(np.float32) * 2 / 255 - 1, atol=1e-6)
    assert np.abs(f[4:]).max() <= 1e-4  # uniform(-1e-4, 1e-4) everywhere else
    m_no_pipe = M.TetrahedraNerf(cfg)
    with pytest.raises(RuntimeError, match="dataparser_scale"):
        m_no_pipe._init_tetrahedra()  # model.py:352-356


def test_image_metrics_and_appearance_embedding_modules():
    """ns-eval / steps_per_eval_image call get_image_metrics_and_images; appearance_embed_dim > 0 must create the embedding and
    widen mlp_head (reference model.py:440-446, 676-713)"""
    import torch

    from tetranerf.nerfstudio import model as M

    cfg = M.TetrahedraNerfConfig(num_tetrahedra_vertices=10, num_tetrahedra_cells=5, appearance_embed_dim=8)
    m = M.TetrahedraNerf(cfg, num_train_data=7)
    assert m.appearance_embedding.weight.shape == (7, 8)
    assert m.mlp_head.layers[0].weight.shape == (128, 128 + 27 + 8)
    assert not m._fused_supported()
    g = torch.Generator().manual_seed(0)
    img = torch.rand((32, 40, 3), generator=g)
    out = {"rgb": (img + 0.05 * torch.rand((32, 40, 3), generator=g)).clamp(0, 1), "accumulation": torch.rand((32, 40, 1), generator=g),
           "depth": 2 + torch.rand((32, 40, 1), generator=g)}
    metrics, images = m.get_image_metrics_and_images(out, {"image": img})
    assert 20 < metrics["psnr"] < 40 and 0.5 < metrics["nerfstudio_ssim"] <= 1.0
    assert images["img"].shape == (32, 80, 3) and images["accumulation"].shape == (32, 40, 3) and images["depth"].shape == (32, 40, 3)
    same, _ = m.get_image_metrics_and_images({**out, "rgb": img}, {"image": img})
    assert same["nerfstudio_ssim"] > 0.9999

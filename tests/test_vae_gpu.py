"""VAE decoder through the C-ABI kernels vs the fp32 oracle restatement (oracle/vae.py), and the row
softmax kernel vs torch.  Tolerance: fp16 storage / fp32 accumulation against fp32 on the same fp16-rounded weights
and latents: rel-L2 <= 5e-3 on the decoded image."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("rows,cols,scale", [(7, 64, 1.0), (33, 1024, 0.5), (16, 16384, 0.044), (5, 32768, 1.0)])
def test_softmax_rows(rows, cols, scale):
    from omg_b200 import ops
    g = torch.Generator().manual_seed(rows + cols)
    x = (torch.randn(rows, cols, generator=g) * 4).half().cuda()
    ref = torch.softmax(x.float() * scale, dim=-1)
    big = torch.zeros(rows, cols + 8, dtype=torch.float16, device="cuda")  # strided rows
    big[:, :cols] = x
    ops.softmax_rows(big[:, :cols], scale)
    got = big[:, :cols].float()
    assert (got - ref).abs().max().item() < 2e-3 * ref.max().item() + 1e-6
    assert torch.allclose(got.sum(-1), torch.ones(rows, device="cuda"), atol=2e-3)
    assert float(big[:, cols:].abs().max()) == 0.0


def test_softmax_rows_rejects_bad_shapes():
    from omg_b200 import ops
    with pytest.raises(RuntimeError):
        ops.softmax_rows(torch.zeros(2, 12, dtype=torch.float16, device="cuda"))
    with pytest.raises(RuntimeError):
        ops.softmax_rows(torch.zeros(2, 16, dtype=torch.float16, device="cuda"), scale=-1.0)


@pytest.mark.parametrize("which,h,w,batch", [("tiny", 16, 16, 2), ("tiny", 8, 24, 1), ("sdxl", 32, 32, 1)])
def test_decode_matches_oracle(which, h, w, batch):
    from omg_b200 import synthetic
    from omg_b200.vae import PackedVaeDecoder, VaeConfig, vae_decoder_param_shapes
    from oracle import vae as ov
    cfg = VaeConfig.tiny() if which == "tiny" else VaeConfig.sdxl()
    ocfg = ov.VaeConfig(block_out_channels=cfg.block_out_channels)
    assert vae_decoder_param_shapes(cfg) == ov.decoder_param_shapes(ocfg)
    sd = {k: v.half().float() for k, v in synthetic.make_vae_state_dict(cfg, seed=3).items()}
    lat = (torch.randn(batch, 4, h, w, generator=torch.Generator().manual_seed(5)) * 0.13025 * 3).half().float()
    torch.set_num_threads(32)
    ref = ov.decode(sd, lat, ocfg)
    dec = PackedVaeDecoder(sd, cfg, device="cuda")
    got = dec.decode(lat.cuda())
    assert got.shape == ref.shape == (batch, 3, 8 * h, 8 * w)
    err = rel(got, ref)
    print(which, h, w, "decode rel-L2", err)
    assert err < 5e-3
    img = dec(lat.cuda(), "pt")
    assert torch.allclose(img.cpu(), ov.postprocess(ref), atol=2e-2)
    assert dec(lat.cuda(), "np").shape == (batch, 8 * h, 8 * w, 3)


def test_pipeline_image_output_uses_the_decoder():
    """output_type='pt' through LoraMultiConceptPipeline with a vae_decoder (lora_pipeline.py:634-661)."""
    from omg_b200 import factory, synthetic
    from omg_b200.config import UNetConfig
    from omg_b200.vae import PackedVaeDecoder, VaeConfig
    wl = factory.build_lora_workload(UNetConfig.tiny(), 128, 2, 8, 6, 7.5)
    wl.pipe.vae_decoder = PackedVaeDecoder(synthetic.make_vae_state_dict(VaeConfig.tiny(), 1), VaeConfig.tiny())
    kw = dict(wl.call_kwargs)
    kw["output_type"] = "pt"
    lat0 = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(0)).half()
    out = wl.pipe(stage=1, latents=lat0, **kw).images
    assert out.shape == (2, 3, 128, 128) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    assert torch.isfinite(out).all()

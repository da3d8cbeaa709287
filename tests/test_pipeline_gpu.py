"""End-to-end parity of the chunk loop (moditalker_b200.pipeline.sample_chunks) on a GPU.

Two arms run the same two-chunk job with the same stand-in autoencoders and the same sampler on the GPU:
  reference-style  the script's own sequence (MToV/sample.py:305-428 with the loader of tools/dataloader_sample.py): pixel work on
                   the CPU (the numpy oracle, itself pinned bit-exactly on the reference's functions), the last frame written to a
                   PNG with cv2 after the BGR swap and read back with PIL + ToTensor, exactly as the script does;
  product          sample_chunks: chunkio kernels, last frame handed over in device memory, files through AsyncFrameWriter.
Every tensor the two arms produce must be bit-identical, and so must the pixels of the files."""
import os

import numpy as np
import pytest
import torch

from conftest import config_by_name
from moditalker_b200 import DDPM, DiffusionWrapper, UNetModel, chunkio
from moditalker_b200.pipeline import Chunk, sample_chunks
from moditalker_b200.synth import synth_state_dict
from oracle import chunkio_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class StandInAE(torch.nn.Module):
    """Deterministic stand-in with the ViTAutoencoder interface of this path (autoencoder_vit.py:212-275): extract
    [B,3,16,256,256] -> [B,4,2048] (xy | yt | xt planes, tanh), decode_from_sample [B,4,2048] -> [(B 16),3,256,256]."""

    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("enc", torch.randn(4, 3, generator=g) * 0.8)
        self.register_buffer("dec", torch.randn(3, 4, generator=g) * 0.9)

    def extract(self, x):
        B = x.shape[0]
        v = torch.nn.functional.avg_pool3d(x, (1, 8, 8))                       # [B,3,16,32,32]
        v = torch.einsum("oc,bcthw->bothw", self.enc, v)
        xy, yt, xt = v.mean(2), v.mean(4), v.mean(3)                           # [B,4,32,32], [B,4,16,32], [B,4,16,32]
        return torch.tanh(torch.cat([xy.reshape(B, 4, -1), yt.reshape(B, 4, -1), xt.reshape(B, 4, -1)], dim=-1))

    def decode_from_sample(self, h):
        B = h.shape[0]
        xy, yt, xt = h[:, :, :1024].view(B, 4, 1, 32, 32), h[:, :, 1024:1536].view(B, 4, 16, 32, 1), h[:, :, 1536:].view(B, 4, 16, 1, 32)
        z = xy + yt + xt                                                       # [B,4,16,32,32]
        img = torch.einsum("oc,bcthw->bothw", self.dec, z).permute(0, 2, 1, 3, 4).reshape(B * 16, 3, 32, 32)
        return 3.0 * torch.tanh(2.0 * torch.nn.functional.interpolate(img, scale_factor=8, mode="nearest"))   # leaves [-1, 1]: the clamp matters


def _job(seed, n_chunks=2, H=90, W=120, N=68):
    rng = np.random.default_rng(seed)
    first = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    chunks = []
    for _ in range(n_chunks):
        kp = rng.uniform(0, H, size=(16, 68, 2))
        kp[:, 33, 1] = rng.uniform(0.3 * H, 0.8 * H, size=16)
        chunks.append(dict(frames=rng.integers(0, 256, size=(16, H, W, 3), dtype=np.uint8), first=first,
                           lm=rng.uniform(-1, 1, size=(16, N, 3)).astype(np.float32), kp=kp))
    return chunks


def _reference_style(ae, ae_l, ddpm, job, k, tmp, ratio):
    """sample.py:305-428 as the script runs it, pixel work on the CPU."""
    import PIL.Image
    from torchvision import transforms

    try:
        import cv2
    except ImportError:
        cv2 = None
    to_tensor = transforms.ToTensor()
    out, prev_dir = [], None
    for it, ch in enumerate(job):
        H, W = ch["frames"].shape[1:3]
        up = lambda a: torch.from_numpy(a)[None].to(DEV)
        x = up(O.prep_frames(ch["frames"], None, 256))
        x_ref = up(O.prep_frames(np.repeat(ch["first"][None], 16, axis=0), None, 256))
        masked_x = up(O.prep_frames(ch["frames"], [O.lower_half_start(H, ch["kp"][t]) for t in range(16)], 256))
        x_l = up(O.rasterize_landmarks(ch["lm"], W))
        z_ = ae.extract(x)
        image_cond_ = ae.extract(x_ref)
        z_l = ae_l.extract(x_l)
        masked_z = ae.extract(masked_x)
        image_cond = image_cond_[:, :, 0:1024]
        if prev_dir is not None:                                                  # sample.py:340-358
            frames_list = []
            for frame in sorted(os.listdir(prev_dir)):
                img = to_tensor(PIL.Image.open(os.path.join(prev_dir, frame))) * 2.0 - 1.0
                frames_list.append(torch.cat([img.unsqueeze(0).to(DEV) for _ in range(16)], dim=0))
            ft = torch.stack(frames_list, dim=0).permute(0, 2, 1, 3, 4).contiguous()   # b t c h w -> b c t h w
            image_cond = ae.extract(ft)[:, :, 0:1024]
        c = torch.cat([z_l, masked_z], dim=1)
        z = ddpm.sample(batch_size=k, cond=c.float(), image_cond=image_cond.float(), noised_start=image_cond_.float(), ratio_=ratio,
                        fix_noise=True)
        fake = ae.decode_from_sample(z).clamp(-1, 1).cpu().numpy()
        frames_u8, last_u8, _ = O.frames_out(fake, k, 16)                         # sample.py:380-399 (pinned on the literal calls)
        prev_dir = os.path.join(tmp, "ref_arm", str(16 * (it + 1)))
        os.makedirs(prev_dir, exist_ok=True)
        for idx in range(k):                                                      # sample.py:388-396: the real file round trip
            lf = ((1 + torch.from_numpy(fake).reshape(k, 16, 3, 256, 256).permute(0, 1, 3, 4, 2)) * 127.5)[idx, -1].numpy()
            if cv2 is not None:
                img = np.rint(cv2.cvtColor(np.asarray(lf, dtype=np.float32), cv2.COLOR_BGR2RGB)).clip(0, 255).astype(np.uint8)
                cv2.imwrite(os.path.join(prev_dir, f"{idx}.png"), img)
            else:
                PIL.Image.fromarray(np.rint(lf).clip(0, 255).astype(np.uint8), "RGB").save(os.path.join(prev_dir, f"{idx}.png"))
        out.append((z.cpu(), frames_u8, last_u8))
    return out


def test_two_chunk_run_is_bit_identical_to_the_scripts_sequence(tmp_path):
    import PIL.Image

    cfg = config_by_name("tiny")
    model = DiffusionWrapper(UNetModel(**cfg))
    model.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True)
    model = model.to(DEV).eval()
    ddpm = DDPM(model, channels=4, image_size=32, sampling_timesteps=12, w=0.0).to(DEV)
    ae, ae_l = StandInAE(1).to(DEV).eval(), StandInAE(2).to(DEV).eval()
    job, k, ratio = _job(5), 1, 0.25

    with torch.no_grad():
        torch.manual_seed(77)
        want = _reference_style(ae, ae_l, ddpm, job, k, str(tmp_path), ratio)
        torch.manual_seed(77)
        chunks = [Chunk(torch.from_numpy(c["frames"]), torch.from_numpy(c["first"]), torch.from_numpy(c["lm"]), c["kp"]) for c in job]
        with chunkio.AsyncFrameWriter() as w:
            got = [(r.z.cpu(), r.frames_u8.cpu().numpy(), r.last_u8.cpu().numpy(), r.next_ref.cpu())
                   for r in sample_chunks(ae, ae_l, ddpm, chunks, device=DEV, batch_size=k, ratio_=ratio, writer=w,
                                          out_dir=str(tmp_path / "prod"))]
    assert len(got) == len(want) == 2
    for it, ((z0, f0, l0), (z1, f1, l1, ref1)) in enumerate(zip(want, got)):
        assert torch.equal(z0, z1), f"latent of chunk {it} differs"       # chunk 1 depends on the chained last frame of chunk 0
        assert np.array_equal(f0, f1) and np.array_equal(l0, l1)
        assert f1.min() == 0 and f1.max() == 255                            # the stand-in decoder overshoots: the clamp was exercised
        # the files of both arms hold the same pixels, and the product's in-memory hand-over equals the PNG read back
        for idx in range(k):
            a = np.asarray(PIL.Image.open(tmp_path / "ref_arm" / str(16 * (it + 1)) / f"{idx}.png"))
            b = np.asarray(PIL.Image.open(tmp_path / "prod" / "references" / str(16 * (it + 1)) / f"{idx}.png"))
            assert np.array_equal(a, b) and np.array_equal(a, l1[idx])
            back = torch.from_numpy(a.astype(np.float32) / np.float32(255)).permute(2, 0, 1) * 2.0 - 1.0
            assert torch.equal(ref1[idx, :, 0], back) and torch.equal(ref1[idx, :, 15], back)
        assert PIL.Image.open(tmp_path / "prod" / "gif" / f"generated_gif_{it}.gif").n_frames == 16
        for t in (0, 15):
            png = np.asarray(PIL.Image.open(tmp_path / "prod" / "frames" / (f"{16 * it + t}".zfill(4) + ".png")))
            assert np.array_equal(png, f1[0, t])
    assert not torch.equal(got[0][0], got[1][0])

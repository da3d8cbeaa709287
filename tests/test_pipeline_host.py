"""Host logic of moditalker_b200.pipeline.sample_chunks on the CPU: which clips reach which model call, how the script's flags
select the start latent and the reference frame, chunk chaining, file names.  The device kernels are replaced by the numpy
oracle (monkeypatched into chunkio for this test only — the product itself never falls back), the autoencoders and the sampler
by recording stand-ins.  The arithmetic is covered on the GPU by tests/test_chunkio_gpu.py and tests/test_pipeline_gpu.py."""
import numpy as np
import pytest
import torch

from moditalker_b200 import chunkio
from moditalker_b200.pipeline import Chunk, sample_chunks
from oracle import chunkio_oracle as O


@pytest.fixture
def cpu_chunkio(monkeypatch):
    def prep(frames, mask_rows=None, resolution=256):
        return torch.from_numpy(O.prep_frames(frames.numpy(), mask_rows, resolution))[None]

    def raster(lm, WH, flip=False):
        return torch.from_numpy(O.rasterize_landmarks(lm.numpy(), WH, flip))[None]

    def fout(dec, batch_size, repeat=16, want_frames=True, want_reference=True):
        return tuple(torch.from_numpy(a) for a in O.frames_out(dec.numpy(), batch_size, repeat))

    monkeypatch.setattr(chunkio, "prep_frames", prep)
    monkeypatch.setattr(chunkio, "rasterize_landmarks", raster)
    monkeypatch.setattr(chunkio, "frames_out", fout)


class RecordingAE:
    def __init__(self, tag, log):
        self.tag, self.log = tag, log

    def extract(self, x):
        self.log.append((self.tag, "extract", x.clone()))
        B = x.shape[0]
        return torch.tanh(x.reshape(B, 3, -1)[:, :, :2048].mean(1, keepdim=True).expand(B, 4, 2048) + len(self.log) * 0.01)

    def decode_from_sample(self, z):
        self.log.append((self.tag, "decode", z.clone()))
        B = z.shape[0]
        return (z[:, :3, :256].reshape(B, 3, 16, 16).repeat_interleave(16, 0).repeat(1, 1, 16, 16) * 1.5).contiguous()


class RecordingSampler:
    def __init__(self, log):
        self.log = log

    def sample(self, **kw):
        self.log.append(("ddpm", "sample", kw))
        return torch.tanh(kw["cond"][:, :4] + kw["image_cond"].mean() + (0 if kw["noised_start"] is None else kw["noised_start"]))


def _chunks(n, H=40, W=48):
    rng = np.random.default_rng(0)
    first = torch.from_numpy(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8))
    out = []
    for _ in range(n):
        kp = rng.uniform(0, H, size=(16, 68, 2))
        out.append(Chunk(torch.from_numpy(rng.integers(0, 256, size=(16, H, W, 3), dtype=np.uint8)), first,
                         torch.from_numpy(rng.uniform(-1, 1, size=(16, 10, 3)).astype(np.float32)), kp))
    return out


def test_clips_flags_and_chaining(cpu_chunkio, tmp_path):
    log = []
    ae, ae_l, ddpm = RecordingAE("rgb", log), RecordingAE("ldmk", log), RecordingSampler(log)
    chunks = _chunks(3)
    with chunkio.AsyncFrameWriter() as w:
        res = list(sample_chunks(ae, ae_l, ddpm, chunks, device="cpu", batch_size=1, ratio_=0.25, writer=w, out_dir=str(tmp_path)))
    assert [r.index for r in res] == [0, 1, 2]
    calls0 = [e for e in log][:6]
    # chunk 0: extract(x), extract(x_ref), ldmk.extract(x_l), extract(masked_x), sample, decode  (sample.py:327-331, 373-381)
    assert [(t, k) for t, k, _ in calls0] == [("rgb", "extract"), ("rgb", "extract"), ("ldmk", "extract"), ("rgb", "extract"),
                                              ("ddpm", "sample"), ("rgb", "decode")]
    x, x_ref, x_l, masked_x = calls0[0][2], calls0[1][2], calls0[2][2], calls0[3][2]
    c0 = chunks[0]
    assert torch.equal(x[0], torch.from_numpy(O.prep_frames(c0.frames_u8.numpy(), None, 256)))
    assert torch.equal(x_ref[0], torch.from_numpy(O.prep_frames(np.repeat(c0.first_frame_u8.numpy()[None], 16, 0), None, 256)))
    rows = [O.lower_half_start(40, c0.keypoints[t]) for t in range(16)]
    assert torch.equal(masked_x[0], torch.from_numpy(O.prep_frames(c0.frames_u8.numpy(), rows, 256)))
    assert torch.equal(x_l[0], torch.from_numpy(O.rasterize_landmarks(c0.landmarks.numpy(), 48)))      # WH = source width
    kw0 = calls0[4][2]
    assert kw0["batch_size"] == 1 and kw0["ratio_"] == 0.25 and kw0["fix_noise"] is True
    assert kw0["cond"].shape == (1, 8, 2048) and kw0["image_cond"].shape == (1, 4, 1024)
    assert kw0["noised_start"] is not None and kw0["noised_start"].shape == (1, 4, 2048)                # --x_noisy_start: extract(x_ref)
    # chunks 1, 2: one more extract — of the previous chunk's last frame as it would be read back from the PNG — feeds image_cond
    later = [e for e in log][6:]
    assert [(t, k) for t, k, _ in later[:7]] == [("rgb", "extract")] * 2 + [("ldmk", "extract")] + [("rgb", "extract")] * 2 + \
        [("ddpm", "sample"), ("rgb", "decode")]
    assert torch.equal(later[4][2], res[0].next_ref)
    assert torch.equal(res[0].next_ref[0, :, 0], (res[0].last_u8[0].float() / 255).permute(2, 0, 1) * 2.0 - 1.0)
    # files under the script's names
    assert sorted(p.name for p in (tmp_path / "gif").iterdir()) == [f"generated_gif_{i}.gif" for i in range(3)]
    assert sorted(p.name for p in (tmp_path / "frames").iterdir()) == [f"{i}".zfill(4) + ".png" for i in range(48)]
    assert sorted(p.name for p in (tmp_path / "references").iterdir()) == ["16", "32", "48"]


def test_flag_variants(cpu_chunkio):
    def run(**kw):
        log = []
        list(sample_chunks(RecordingAE("rgb", log), RecordingAE("ldmk", log), RecordingSampler(log), _chunks(2), device="cpu", **kw))
        return log, [e[2] for e in log if e[1] == "sample"]

    log, s = run(use_last_as_reference=False)
    assert sum(1 for e in log if e[:2] == ("rgb", "extract")) == 6                    # no extract of a chained frame
    log, s = run(x_noisy_start=False)
    assert all(k["noised_start"] is None for k in s)                                  # plain DDIM from noise
    log, s = run(x_noisy_start=False, refvid_noisy_start=True)
    assert all(k["noised_start"] is not None for k in s)                              # extract(x) as the start latent
    first_extract = [e[2] for e in log if e[:2] == ("rgb", "extract")][0]
    log2, s2 = run(batch_size=2)
    assert s2[0]["batch_size"] == 2 and s2[0]["cond"].shape[0] == 2 and s2[0]["image_cond"].shape[0] == 2
    assert torch.equal([e[2] for e in log2 if e[:2] == ("rgb", "extract")][0][1], first_extract[0])   # the clips are shared by the k samples

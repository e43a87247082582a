"""GPU parity of the fused LFCC kernel (through the C ABI) against the CPU oracle and the
golden vectors generated from the reference.  Tolerance: |a-b| <= 1e-4 (|b| + 1) (tolerances.py);
frame counts, frame indexing and crop/pad maps are exact."""
import os

import numpy as np
import pytest
import torch

from oracle import lfcc_oracle as lo
from oracle import state_spec as ss
from tolerances import lfcc_close, lfcc_worst

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["tc", "fft"])
def mod(request):
    """Both device implementations: "tc" = tensor-core folded DFT (csrc/lfcc_tc.cu, the default), "fft" = radix FFT on
    CUDA cores (csrc/lfcc.cu)."""
    from asvspoof2021_air_b200.feature_extraction import LFCC
    m = LFCC(320, 160, 512, 16000, 20).cuda()
    m.impl = request.param
    return m


@pytest.fixture(scope="module")
def gl(golden_dir):
    return np.load(os.path.join(golden_dir, "lfcc_golden.npz"))


def test_config0_32_waves_vs_reference_golden_and_oracle(mod, gl):
    w = ss.seeded_waves(32, 64000, seed=0, edge_rows=True)
    keep = w.clone()
    y = mod(w.cuda())
    assert y.shape == (35, 401, 60) and y.dtype == torch.float32
    y = y.cpu().numpy()
    assert torch.equal(w, keep)                      # input not modified
    rows = gl["full_rows"]
    assert lfcc_close(y[rows], gl["full"]).all(), lfcc_worst(y[rows], gl["full"])
    assert lfcc_close(y[:, gl["sparse_frames"]], gl["sparse"]).all()
    assert (np.abs(y.astype(np.float64).sum(axis=1) - gl["colsum"]) <= 1e-4 * (np.abs(gl["colsum"]) + 401.0)).all()
    o = lo.lfcc(w.numpy())
    assert lfcc_close(y, o).all(), lfcc_worst(y, o)
    print("worst LFCC deviation vs oracle: %.3g (tolerance 1e-4)" % lfcc_worst(y, o))


@pytest.mark.parametrize("L", [3200, 12345, 5920, 6079, 321, 800, 160, 159, 1])
def test_ragged_lengths(mod, gl, L):
    w = ss.seeded_waves(2, L, seed=L)
    y = mod(w.cuda()).cpu().numpy()
    assert y.shape == (2, 1 + L // 160, 60)
    o = lo.lfcc(w.numpy())
    assert lfcc_close(y, o).all(), lfcc_worst(y, o)
    key = "ragged_%d" % L
    if key in gl:
        assert lfcc_close(y, gl[key]).all()


def test_lengths_vector_matches_per_utterance_runs(mod):
    lens = [64000, 40000, 12345, 321, 63999]
    w = ss.seeded_waves(len(lens), 64000, seed=5)
    y = mod.extract(w.cuda(), lengths=torch.tensor(lens), feat_len=0, layout="btd", dtype=torch.float32).cpu().numpy()
    for b, n in enumerate(lens):
        T = 1 + n // 160
        o = lo.lfcc(w[b:b + 1, :n].numpy())[0]
        assert lfcc_close(y[b, :T], o).all(), (b, lfcc_worst(y[b, :T], o))
        assert (y[b, T:] == 0).all()


@pytest.mark.parametrize("padding", ["repeat", "zero", "silence"])
def test_pad_policies_are_index_exact(mod, padding):
    w = ss.seeded_waves(3, 64000, seed=9)
    base = mod(w.cuda())                                             # (3,401,60)
    out = mod.extract(w.cuda(), feat_len=750, padding=padding, layout="btd", dtype=torch.float32)
    fmap = lo.frame_index_map(401, 750, padding)
    sil = mod.silence_vector(base.device).cpu().numpy()
    want = lo.apply_frame_map(base.cpu().numpy(), fmap, silence=sil)
    assert np.array_equal(out.cpu().numpy(), want)                   # pure index work: bit exact
    assert lfcc_close(sil, lo.silence_vector()).all()


def test_crop_policy_is_index_exact(mod):
    L = 160 * 999
    w = ss.seeded_waves(2, L, seed=4)
    base = mod(w.cuda()).cpu().numpy()                               # T = 1000
    start = torch.tensor([17, 249])
    out = mod.extract(w.cuda(), feat_len=750, padding="repeat", start=start, layout="btd",
                      dtype=torch.float32).cpu().numpy()
    for b in range(2):
        assert np.array_equal(out[b], base[b, lo.frame_index_map(1000, 750, "repeat", int(start[b]))])


def test_model_layouts(mod):
    w = ss.seeded_waves(2, 64000, seed=2)
    btd = mod.extract(w.cuda(), feat_len=750, padding="repeat", layout="btd", dtype=torch.float32)
    r = mod.extract(w.cuda(), feat_len=750, padding="repeat", layout="resnet", dtype=torch.float32)
    assert r.shape == (2, 1, 60, 750)
    assert torch.equal(r[:, 0], btd.transpose(1, 2))                 # main_train.py:338
    e = mod.extract(w.cuda(), feat_len=750, padding="repeat", layout="ecapa", dtype=torch.bfloat16)
    assert e.shape == (2, 750, 64) and (e[:, :, 60:] == 0).all()
    assert torch.equal(e[:, :, :60], btd.to(torch.bfloat16))


def test_full_size_properties(mod):
    """BASELINE config sizes (B=256): size-independent properties instead of a CPU oracle run."""
    w = ss.seeded_waves(256, 64000, seed=1).cuda()
    y = mod(w)
    assert y.shape == (256, 401, 60) and torch.isfinite(y).all()
    # delta linearity / structure: delta[t] = c[t+1] - c[t-1] with replicate edges
    c, d, dd = y[..., :20], y[..., 20:40], y[..., 40:]
    cp = torch.cat([c[:, :1], c, c[:, -1:]], 1)
    assert torch.allclose(d, cp[:, 2:] - cp[:, :-2], atol=1e-5)
    dp = torch.cat([d[:, :1], d, d[:, -1:]], 1)
    assert torch.allclose(dd, dp[:, 2:] - dp[:, :-2], atol=2e-5)
    # batch independence + determinism: rows 0..31 equal a B=32 run bit for bit
    assert torch.equal(mod(w[:32].clone()), y[:32])
    # scaling the wave by 2 adds log10(4) to every filterbank log-energy -> only c0 moves
    y2 = mod(w[:8] * 2.0)
    shift = np.log10(4.0) * np.sqrt(20.0)            # ortho DCT: c0 = sum(fbe)/sqrt(20)
    assert torch.allclose(y2[..., 0] - y[:8, :, 0], torch.full_like(y2[..., 0], shift), atol=1e-3)
    assert torch.allclose(y2[..., 1:20], y[:8, :, 1:20], atol=1e-3)


@pytest.mark.parametrize("scale", [1.0, 1e-3])
def test_tensor_core_kernel_holds_the_fp32_bar_on_high_dynamic_range_speech(scale):
    """The 3-term split of csrc/lfcc_tc.cu (fp16 hi parts, bf16 residual of x, fp16 residual of w) on a speech-like wave
    whose weak bands lie ~60 dB under its strongest harmonics, loud and 60 dB quieter: within 5e-5 (|ref|+1) of the float64
    oracle (tolerance of the contract 1e-4; a bf16 / bf16 split gives 1.0e-4 here, an fp16 / fp16 split 3e-4 on the quiet one;
    tests/test_oracle.py pins the same numbers on the CPU emulation)."""
    from asvspoof2021_air_b200.feature_extraction import LFCC
    from tolerances import lfcc_worst
    n = np.arange(16000)
    rng = np.random.RandomState(0)
    speech = sum(0.3 / h * np.sin(2 * np.pi * 140 * h * n / 16000) for h in range(1, 9)) + 3e-4 * rng.randn(16000)
    w = (scale * speech)[None].astype(np.float32)
    want = lo.lfcc(w)
    m = LFCC(320, 160, 512, 16000, 20).cuda()
    assert m.impl_for(torch.float32) in ("tc", "fft")
    m.impl = "tc"
    got = m(torch.from_numpy(w).cuda()).cpu().numpy()
    assert got.shape == want.shape
    assert lfcc_worst(got[:, :, :20], want[:, :, :20]) < 5e-5, lfcc_worst(got[:, :, :20], want[:, :, :20])
    assert lfcc_worst(got, want) < 1e-4, lfcc_worst(got, want)

"""torch fp32 CPU restatement of the reference LFCC forward with the modern torch.stft signature.

TEST INFRASTRUCTURE ONLY (see oracle/lfcc_oracle.py).  This is the arithmetic the reference runs
(feature_extraction.py:93-138: same torch ops, fp32, MKL FFT) and is what bench.py times as the
CPU baseline / `--impl reference` arm on the GPU box, where /root/reference is not mounted.
Pinned against the golden vectors in tests/test_oracle.py.
"""
import torch
import torch.nn.functional as F

from . import lfcc_oracle as lo


class TorchLFCC:
    def __init__(self, fl=320, fs=160, fn=512, sr=16000, filter_num=20):
        self.fl, self.fs, self.fn = fl, fs, fn
        self.fb = torch.from_numpy(lo.linear_filterbank(fn, sr, filter_num))
        self.dct = torch.from_numpy(lo.dct_matrix(filter_num)).float()
        self.win = torch.hamming_window(fl)

    @staticmethod
    def _delta(x):                                               # feature_extraction.py:41-58
        xp = F.pad(x.unsqueeze(1), (0, 0, 1, 1), "replicate").squeeze(1)
        return xp[:, 2:] - xp[:, :-2]

    def __call__(self, x):
        x = x.clone()
        x[:, 1:] = x[:, 1:] - 0.97 * x[:, 0:-1]                  # :105-106
        st = torch.stft(x, self.fn, self.fs, self.fl, window=self.win, onesided=True,
                        pad_mode="constant", return_complex=True)      # :109-111
        sp = (st.real ** 2 + st.imag ** 2).permute(0, 2, 1).contiguous()   # :113
        fbe = torch.log10(torch.matmul(sp, self.fb) + torch.finfo(torch.float32).eps)   # :116-117
        c = F.linear(fbe, self.dct)                               # :120
        d = self._delta(c)
        return torch.cat((c, d, self._delta(d)), 2)               # :130-133

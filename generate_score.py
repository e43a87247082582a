#!/usr/bin/env python
"""Scoring driver with the argparse surface of the reference's generate_score.py:10-36, batched on the fused
B200 path: raw waves -> on-device LFCC / pad -> eval-mode model -> OC-Softmax score.

Score lines follow generate_score.py:115-119: `"%s %s\\n" % (utt, +cos(feat, centre))` (higher = bonafide) and, for
the `19*` tasks, a third `bonafide` / `spoof` column.  The reference scores with batch size 1 in a Python loop;
here utterances are scored in batches (--batch_size, default 1024).  Model pickles: whole-module files as
written by main_train.py (`anti-spoofing_feat_model.pt`; the reference's loader expects the older name
`anti-spoofing_cqcc_model.pt`, generate_score.py:135 -- both are accepted) + `anti-spoofing_loss_model.pt`.
New flags: --synthetic N / --wave_dir / --protocol (the reference's hard-coded dataset classes read
pre-extracted .pt features from the authors' disks, generate_score.py:50-70).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

TASKS = ["LA", "DF", "19dev", "19laaugdev", "19lapaaugdev", "19dfaugdev", "19dfpaaugdev", "19eval"]


def build_parser():
    p = argparse.ArgumentParser("load model scores")
    p.add_argument("--model_folder", type=str, default="/data/xinhui/models/")
    p.add_argument("-n", "--model_name", type=str, required=True, default="lfcc_ecapa512ctst_ocs")
    p.add_argument("-s", "--score_dir", type=str, default="/data/neil/scores")
    p.add_argument("-t", "--task", type=str, required=True, default="LA", choices=TASKS)
    p.add_argument("-l", "--loss", default=None, required=False, choices=[None, "ocsoftmax", "amsoftmax", "p2sgrad"])
    p.add_argument("--gpu", type=str, default="0")
    new = p.add_argument_group("fused raw-wave path (not in the reference)")
    new.add_argument("--synthetic", type=int, default=0)
    new.add_argument("--wave_dir", type=str, default=None)
    new.add_argument("--protocol", type=str, default=None)
    new.add_argument("--packed_waves", type=str, default=None, help="prefix of a corpus written by asvspoof2021_air_b200.data pack")
    new.add_argument("--batch_size", type=int, default=1024)
    new.add_argument("--feat_len", type=int, default=750)
    new.add_argument("--padding", type=str, default="repeat", choices=["zero", "repeat", "silence"])
    new.add_argument("--attention_noise", action="store_true",
                     help="resnet: add SelfAttention's 1e-5 * randn (resnet.py:38-42) as the reference does even when "
                          "scoring; off by default so that score files are reproducible")
    return p


def init(argv=None):
    args = build_parser().parse_args(argv)
    if "LOCAL_RANK" not in os.environ:
        os.environ["CUDA_VISIBLE_DEVICES"] = args.gpu
    args.cuda = torch.cuda.is_available()
    args.device = torch.device("cuda" if args.cuda else "cpu")
    args.out_score_dir = "./scores" if "19" in args.task else args.score_dir      # generate_score.py:31-34
    return args


def score_file_path(output_score_path, model_name, task):
    """generate_score.py:76-82."""
    if "19" in task:
        return os.path.join(output_score_path, model_name + "_" + task + "_score.txt")
    d = os.path.join(output_score_path, model_name + "_" + task)
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "score.txt")


def format_line(task, utt, score, label):
    """generate_score.py:113-119 (score is +cos; label 1 = spoof)."""
    if "19" in task:
        return "%s %s %s\n" % (utt, score, "spoof" if label else "bonafide")
    return "%s %s\n" % (utt, score)


def _device():
    return torch.device("cuda", torch.cuda.current_device())


MAX_BATCH = 16384        # air_ocsoftmax_fwd_bwd: (D + 3B + 32) floats of shared memory, 200 KB limit at D = 256


def test_on_ASVspoof2021(task, feat_model_path, loss_model_path, output_score_path, model_name, add_loss, args):
    from asvspoof2021_air_b200 import data
    from asvspoof2021_air_b200.trainer import Trainer
    if add_loss != "ocsoftmax":
        raise SystemExit("only -l ocsoftmax is implemented on the fused path (SURVEY.md section 2.1)")
    if not torch.cuda.is_available():
        raise SystemExit("generate_score.py needs a CUDA device: the fused path has no CPU fallback")
    if not 1 <= args.batch_size <= MAX_BATCH:
        raise SystemExit("--batch_size must be in [1, %d] (one OC-Softmax launch holds the whole batch on one SM)" % MAX_BATCH)
    from asvspoof2021_air_b200 import compat
    model = compat.load_module(feat_model_path)              # this package's pickles and the reference's own
    loss_model = compat.load_module(loss_model_path)
    arch = "ecapa" if type(model).__name__ == "Res2Net2" else "resnet"
    tr = Trainer(arch=arch, enc_dim=loss_model.center.shape[1], feat_len=args.feat_len, padding=args.padding,
                 r_real=loss_model.r_real, r_fake=loss_model.r_fake, alpha=loss_model.alpha, device="cuda",
                 attention_noise=args.attention_noise)
    tr.load_modules(model, loss_model)
    if args.packed_waves:
        src = data.PackedWaves(args.packed_waves, args.feat_len)
    elif args.wave_dir:
        src = data.WaveFolder(args.wave_dir, args.protocol, args.feat_len)
    elif args.synthetic > 0:
        src = data.SyntheticWaves(args.synthetic, feat_len=args.feat_len)
    else:
        raise SystemExit("no input: pass --wave_dir/--protocol or --synthetic N")
    os.makedirs(output_score_path, exist_ok=True)
    path = score_file_path(output_score_path, model_name, task)
    order = [list(range(lo, min(lo + args.batch_size, len(src)))) for lo in range(0, len(src), args.batch_size)]
    with open(path, "w") as f:
        # decode + H2D of the next batches overlap the forward pass of this one (data.Prefetcher)
        for batch in data.Prefetcher(src, order, depth=2, device=_device()):
            waves, lengths, _, names, start = batch
            s = tr.score_step(waves, lengths, start).cpu()
            for j, name in enumerate(names):
                f.write(format_line(task, name, float(s[j]), int(batch.labels_host[j])))
    return path


if __name__ == "__main__":
    args = init()
    model_dir = os.path.join(args.model_folder, args.model_name)
    model_path = os.path.join(model_dir, "anti-spoofing_feat_model.pt")
    if not os.path.exists(model_path):
        model_path = os.path.join(model_dir, "anti-spoofing_cqcc_model.pt")     # generate_score.py:135
    loss_model_path = os.path.join(model_dir, "anti-spoofing_loss_model.pt")
    print(test_on_ASVspoof2021(args.task, model_path, loss_model_path, args.out_score_dir, args.model_name, args.loss, args))

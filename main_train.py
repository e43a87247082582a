#!/usr/bin/env python
"""Training driver with the argparse surface of the reference's main_train.py:23-95 (every flag, same
names and defaults), running the fused B200 path: raw waves -> on-device LFCC / crop / pad -> ResNet-18 or
ECAPA-TDNN -> OC-Softmax (`--add_loss ang_iso`) -> backward -> Adam(L2) + SGD(centre), one process per GPU
under torchrun with an NCCL gradient all-reduce.

Kept from the reference: args.json, train_loss.log (`epoch\\tstep\\tloss`), dev_loss.log, per-epoch whole-module
checkpoints `checkpoint/anti-spoofing_{feat,loss}_model_%d.pt`, best-dev copies `anti-spoofing_{feat,loss}_model.pt`,
lr * lr_decay^(epoch // interval), --continue_training (model + loss module only, main_train.py:172-173).
New flags: --synthetic N (seeded synthetic utterances per epoch), --wave_dir / --protocol (+ --dev_*) for raw
audio, --steps_per_epoch, --log_every.  Out of the hot-path scope and rejected at run time with a clear
message: models other than resnet / ecapa, --add_loss other than ang_iso, --visualize.
--ADV_AUG (main_train.py:211-224,377-453) runs from raw waves: --wave_dir + --aug_wave_dir + --protocol, or a packed
corpus of them.  Deviation: validation uses the plain dev source (--dev_*); the reference validates on the augmented
dev set and logs the channel accuracy there (main_train.py:489-577), which needs the augmented dev audio.
"""
import argparse
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from asvspoof2021_air_b200.utils import setup_seed, str2bool  # noqa: E402

# (flags, kwargs) -- names / defaults / choices as in main_train.py:26-93
_REFERENCE_FLAGS = [
    (("--seed",), dict(type=int, default=688)),
    (("-a", "--access_type"), dict(type=str, default="LA")),
    (("-d", "--path_to_database"), dict(type=str, default="/data/neil/DS_10283_3336/")),
    (("-f", "--path_to_features"), dict(type=str, default="/data2/neil/ASVspoof2019LA/")),
    (("-o", "--out_fold"), dict(type=str, required=True, default="./models/try/")),
    (("--ratio",), dict(type=float, default=0.5)),
    (("--feat",), dict(type=str, default="LFCC", choices=["CQCC", "LFCC"])),
    (("--feat_len",), dict(type=int, default=750)),
    (("--pad_chop",), dict(type=str2bool, nargs="?", const=True, default=True)),
    (("--padding",), dict(type=str, default="repeat", choices=["zero", "repeat", "silence"])),
    (("--enc_dim",), dict(type=int, default=256)),
    (("-m", "--model"), dict(default="lcnn", choices=["cnn", "resnet", "lcnn", "res2net", "ecapa"])),
    (("--num_epochs",), dict(type=int, default=200)),
    (("--batch_size",), dict(type=int, default=64)),
    (("--lr",), dict(type=float, default=0.0005)),
    (("--lr_decay",), dict(type=float, default=0.5)),
    (("--interval",), dict(type=int, default=30)),
    (("--beta_1",), dict(type=float, default=0.9)),
    (("--beta_2",), dict(type=float, default=0.999)),
    (("--eps",), dict(type=float, default=1e-8)),
    (("--gpu",), dict(type=str, default="1")),
    (("--num_workers",), dict(type=int, default=0)),
    (("--base_loss",), dict(type=str, default="ce", choices=["ce", "bce"])),
    (("--add_loss",), dict(type=str, default=None, choices=[None, "isolate", "ang_iso", "p2sgrad"])),
    (("--weight_loss",), dict(type=float, default=1)),
    (("--r_real",), dict(type=float, default=0.9)),
    (("--r_fake",), dict(type=float, default=0.2)),
    (("--alpha",), dict(type=float, default=20)),
    (("--num_centers",), dict(type=int, default=3)),
    (("--visualize",), dict(action="store_true")),
    (("--test_only",), dict(action="store_true")),
    (("--continue_training",), dict(action="store_true")),
    (("--ADV_AUG",), dict(type=str2bool, nargs="?", const=True, default=False)),
    (("--LA_aug",), dict(type=str2bool, nargs="?", const=True, default=False)),
    (("--DF_aug",), dict(type=str2bool, nargs="?", const=True, default=False)),
    (("--LAPA_aug",), dict(type=str2bool, nargs="?", const=True, default=False)),
    (("--DFPA_aug",), dict(type=str2bool, nargs="?", const=True, default=False)),
    (("--lambda_",), dict(type=float, default=0.05)),
    (("--lr_d",), dict(type=float, default=0.0001)),
    (("--pre_train",), dict(action="store_true")),
    (("--test_on_eval",), dict(action="store_true")),
]


def build_parser():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    for flags, kw in _REFERENCE_FLAGS:
        parser.add_argument(*flags, **kw)
    new = parser.add_argument_group("fused raw-wave path (not in the reference)")
    new.add_argument("--synthetic", type=int, default=0, help="train on N seeded synthetic 4 s utterances per epoch")
    new.add_argument("--dev_synthetic", type=int, default=0, help="validate on N synthetic utterances")
    new.add_argument("--wave_dir", type=str, default=None, help="folder of .wav / .npy training waves")
    new.add_argument("--protocol", type=str, default=None, help="protocol file for --wave_dir")
    new.add_argument("--aug_wave_dir", type=str, default=None,
                     help="--ADV_AUG: folder of <utt>_<channel>[_<device>] augmented copies (raw_dataset.py:149-300)")
    new.add_argument("--packed_waves", type=str, default=None,
                     help="prefix of a corpus written by `python -m asvspoof2021_air_b200.data pack` (decode FLAC once)")
    new.add_argument("--dev_packed_waves", type=str, default=None)
    new.add_argument("--dev_wave_dir", type=str, default=None)
    new.add_argument("--dev_protocol", type=str, default=None)
    new.add_argument("--steps_per_epoch", type=int, default=0, help="0: one pass over the source")
    new.add_argument("--log_every", type=int, default=50, help="steps between host reads of the device losses")
    new.add_argument("--attention_noise", type=str2bool, nargs="?", const=True, default=True,
                     help="resnet: the 1e-5 * randn of SelfAttention's pooled std (resnet.py:38-42), drawn on device")
    return parser


def init_params(argv=None):
    args = build_parser().parse_args(argv)
    assert 0 < args.ratio <= 1                                           # main_train.py:98
    rank = int(os.environ.get("RANK", "0"))
    if "LOCAL_RANK" not in os.environ:
        os.environ["CUDA_VISIBLE_DEVICES"] = args.gpu                     # main_train.py:101 (single-process launch)
    setup_seed(args.seed)
    if not (args.test_only or args.continue_training) and rank == 0:
        if os.path.exists(args.out_fold):
            shutil.rmtree(args.out_fold)
        os.makedirs(os.path.join(args.out_fold, "checkpoint"))
        with open(os.path.join(args.out_fold, "args.json"), "w") as f:
            f.write(json.dumps(vars(args), sort_keys=True, separators=("\n", ":")))
        for name, what in (("train_loss.log", "training"), ("dev_loss.log", "validation"), ("test_loss.log", "test")):
            with open(os.path.join(args.out_fold, name), "w") as f:
                f.write("Start recording %s loss ...\n" % what)
    args.cuda = torch.cuda.is_available()
    args.device = torch.device("cuda" if args.cuda else "cpu")
    return args


def adjust_learning_rate(args, lr, epoch_num):
    """main_train.py:144-147."""
    return lr * (args.lr_decay ** (epoch_num // args.interval))


def _reject_out_of_scope(args):
    if args.model not in ("resnet", "ecapa"):
        raise SystemExit("--model %s is outside the B200 hot path (resnet / ecapa only, SURVEY.md section 2.1)" % args.model)
    if args.add_loss != "ang_iso":
        raise SystemExit("only --add_loss ang_iso (OC-Softmax) is implemented on the fused path")
    if args.visualize:
        raise SystemExit("--visualize is outside the B200 hot path")
    if args.ADV_AUG:
        if sum(bool(x) for x in (args.LA_aug, args.DF_aug, args.LAPA_aug, args.DFPA_aug)) != 1:
            raise SystemExit("--ADV_AUG needs exactly one of --LA_aug / --DF_aug / --LAPA_aug / --DFPA_aug (main_train.py:212)")
        if not ((args.wave_dir and args.aug_wave_dir and args.protocol) or args.packed_waves):
            raise SystemExit("--ADV_AUG trains from raw waves: pass --wave_dir (originals), --aug_wave_dir and --protocol, "
                             "or --packed_waves of a corpus packed from them")
    if args.feat != "LFCC":
        raise SystemExit("only --feat LFCC is implemented (computed on device from raw waves)")


def _device():
    return torch.device("cuda", torch.cuda.current_device())


def dev_eer(score_chunks, label_chunks, world):
    """EER of the validation scores as the reference computes it (main_train.py:660-664: label 0 = bona fide is the
    target class; the better of the two score orientations), sorted and reduced on the GPU (csrc/det.cu).  Under
    data parallelism every rank contributes its shard of the dev set (one padded all-gather)."""
    from asvspoof2021_air_b200 import eval_metrics as em
    if not score_chunks:
        scores, labels = torch.empty(0, device=_device()), torch.empty(0, dtype=torch.long, device=_device())
    else:
        scores, labels = torch.cat(score_chunks).float(), torch.cat(label_chunks).long()
    if world > 1:
        from asvspoof2021_air_b200 import parallel
        scores, labels = parallel.gather_ragged([scores, labels])
    tar, non = scores[labels == 0], scores[labels == 1]
    if tar.numel() == 0 or non.numel() == 0:
        return float("nan")
    return min(em.det(tar, non).host()["eer"], em.det(tar, non, negate=True).host()["eer"])


def _source(args, dev=False):
    from asvspoof2021_air_b200 import data
    n = args.dev_synthetic if dev else args.synthetic
    folder = args.dev_wave_dir if dev else args.wave_dir
    packed = args.dev_packed_waves if dev else args.packed_waves
    if packed:
        return data.PackedWaves(packed, args.feat_len, args.seed + (1 if dev else 0))
    if folder and args.ADV_AUG and not dev:
        kind = "LA" if args.LA_aug else "DF" if args.DF_aug else "LAPA" if args.LAPA_aug else "DFPA"
        return data.AugWaveFolder(folder, args.aug_wave_dir, args.protocol, kind, args.feat_len, args.seed)
    if folder:
        return data.WaveFolder(folder, args.dev_protocol if dev else args.protocol, args.feat_len, args.seed + (1 if dev else 0))
    if n > 0:
        return data.SyntheticWaves(n, seed=args.seed + (7 if dev else 0), feat_len=args.feat_len)
    if dev:
        return None
    raise SystemExit("no training data: the reference's pre-extracted .pt feature folders (--path_to_features) are replaced "
                     "by on-device LFCC; pass --wave_dir/--protocol or --synthetic N")


def train(args):
    from asvspoof2021_air_b200 import parallel
    from asvspoof2021_air_b200.trainer import Trainer
    _reject_out_of_scope(args)
    if not torch.cuda.is_available():
        raise SystemExit("main_train.py needs a CUDA device: the fused path has no CPU fallback")
    rank, world, _ = parallel.init_from_env()
    pg = torch.distributed.group.WORLD if world > 1 else None
    tr = Trainer(arch=args.model, enc_dim=args.enc_dim, feat_len=args.feat_len, padding=args.padding, lr=args.lr,
                 beta_1=args.beta_1, beta_2=args.beta_2, eps=args.eps, weight_decay=0.0005, r_real=args.r_real,
                 r_fake=args.r_fake, alpha=args.alpha, weight_loss=args.weight_loss, device="cuda", process_group=pg,
                 seed=args.seed, attention_noise=args.attention_noise)
    if args.continue_training:                                            # main_train.py:172-173
        from asvspoof2021_air_b200 import compat
        model = compat.load_module(os.path.join(args.out_fold, "anti-spoofing_feat_model.pt"))
        lp = os.path.join(args.out_fold, "anti-spoofing_loss_model.pt")
        tr.load_modules(model, compat.load_module(lp) if os.path.exists(lp) else None)
    from asvspoof2021_air_b200 import data
    device = _device()
    src, dev_src = _source(args), _source(args, dev=True)
    per_rank = args.batch_size
    order_rng = np.random.RandomState(args.seed)
    steps = args.steps_per_epoch or max(1, len(src) // (per_rank * world))
    adv = bool(args.ADV_AUG)
    if adv:                                                               # main_train.py:211-224
        if not getattr(src, "channel_names", None):
            raise SystemExit("--ADV_AUG: the training source carries no channel labels (a corpus packed without "
                             "--aug_wave_dir?); pack originals + augmented copies together or pass --wave_dir/--aug_wave_dir")
        heads = [len(src.channel_names)] + ([len(src.device_names)] if src.device_names else [])
        tr.attach_adversaries(heads, lambda_=args.lambda_, lr_d=args.lr_d, seed=args.seed)
        steps = args.steps_per_epoch or max(1, src.n_ori // max(1, int(per_rank * args.ratio) * world))
        aug_rng = np.random.RandomState(args.seed + 1000 * rank)
    prev_loss, early_stop_cnt = 1e8, 0
    feat_model = loss_model = None
    for epoch in range(args.num_epochs):
        lr = adjust_learning_rate(args, args.lr, epoch)
        pending = []
        if adv:                                                           # two half-batches per step, main_train.py:226-233,309-325
            tr.lr_d = adjust_learning_rate(args, args.lr_d, epoch)
            order = data.half_batches(src.n_ori, len(src), per_rank, args.ratio, steps, aug_rng)
            seen_m = seen_c = 0
            right_m = torch.zeros(1, dtype=torch.int64, device=device)
            right_c = torch.zeros(1, dtype=torch.int64, device=device)
        else:
            perm = order_rng.permutation(len(src))                        # SubsetRandomSampler, main_train.py:226-242
            order = [[perm[((step * world + rank) * per_rank + j) % len(src)] for j in range(per_rank)] for step in range(steps)]
        # decode / collate / H2D of the next batches run on a host thread + copy stream while this step computes
        for step, batch in enumerate(data.Prefetcher(src, order, depth=2, device=device)):
            waves, lengths, labels, _, start = batch
            # main_train.py:377 / :420: the gradient-reversed term joins after epoch 0; the classifier's own step (second
            # forward on detached features) runs in every epoch, epoch 0 included
            adv_now = adv and epoch > 0
            loss = tr.train_step(waves, labels, lengths=lengths, start=start, lr=lr,
                                 channels=batch.channels if adv else None, step_seed=epoch * steps + step, grl=adv_now)
            rec = [step, loss.clone()]
            if adv:                                                       # main_train.py:428-429 (every epoch)
                right_c += tr.adv_stats_c[0][1]
                seen_c += waves.shape[0]
            if adv_now:                                                   # main_train.py:383-384,471-477
                right_m += tr.adv_stats[0][1]
                seen_m += waves.shape[0]
                rec += [sum(s[0] for s in tr.adv_stats).clone(), right_m.clone(), seen_m, right_c.clone(), seen_c]
            pending.append(rec)
            if len(pending) >= args.log_every or step == steps - 1:
                if rank == 0:                                             # main_train.py:471-481, batched host reads
                    with open(os.path.join(args.out_fold, "train_loss.log"), "a") as log:
                        for r in pending:
                            if len(r) == 2:
                                log.write("%d\t%d\t%s\n" % (epoch, r[0], float(r[1])))
                            else:
                                log.write("%d\t%d\t%s\t%s\t%s\t%s\n" % (epoch, r[0], float(r[2]), 100.0 * int(r[3]) / r[4],
                                                                         100.0 * int(r[5]) / r[6], float(r[1])))
                pending = []
        val = float("nan")
        if dev_src is not None:
            tot, cnt, dev_scores, dev_labels = 0.0, 0, [], []
            dev_order = [list(range(lo, min(lo + per_rank, len(dev_src))))
                         for lo in range(rank * per_rank, len(dev_src), per_rank * world)]
            for waves, lengths, labels, names, start in data.Prefetcher(dev_src, dev_order, depth=2, device=device):
                l, sc = tr.eval_loss(waves, labels, lengths, start)
                tot, cnt = tot + float(l) * len(names), cnt + len(names)
                dev_scores.append(sc.clone())
                dev_labels.append(labels)
            t = torch.tensor([tot, cnt], device=device, dtype=torch.float64)
            if world > 1:
                torch.distributed.all_reduce(t)
            val = float(t[0] / t[1].clamp(min=1))
            eer = dev_eer(dev_scores, dev_labels, world)
            if rank == 0:                                                 # main_train.py:598-601 / :666-667
                with open(os.path.join(args.out_fold, "dev_loss.log"), "a") as log:
                    log.write("%d\t%s\t%s\n" % (epoch, val, eer))
        if rank == 0:                                                     # main_train.py:674-706
            feat_model, loss_model = tr.modules()
            ck = os.path.join(args.out_fold, "checkpoint")
            torch.save(feat_model, os.path.join(ck, "anti-spoofing_feat_model_%d.pt" % (epoch + 1)))
            torch.save(loss_model, os.path.join(ck, "anti-spoofing_loss_model_%d.pt" % (epoch + 1)))
            if not (val >= prev_loss):                                    # also when there is no dev set (nan)
                torch.save(feat_model, os.path.join(args.out_fold, "anti-spoofing_feat_model.pt"))
                torch.save(loss_model, os.path.join(args.out_fold, "anti-spoofing_loss_model.pt"))
        if val < prev_loss:
            prev_loss, early_stop_cnt = val, 0
        elif val == val:
            early_stop_cnt += 1
        if early_stop_cnt == 500:                                         # main_train.py:711-714
            if rank == 0:
                with open(os.path.join(args.out_fold, "args.json"), "a") as f:
                    f.write("\nTrained Epochs: %d\n" % (epoch - 499))
            break
    if world > 1:
        torch.distributed.destroy_process_group()
    return feat_model, loss_model


if __name__ == "__main__":
    train(init_params())

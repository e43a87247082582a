"""The fused train / score step: raw waves in, optimiser step (or scores) out, all on device.

Mirrors the step body of main_train.py:310-418 for `--add_loss ang_iso` (OC-Softmax) --
LFCC (fused, replaces preprocess.py + dataset.py crop/pad) -> model -> CE (logged) -> OC-Softmax ->
backward -> Adam(L2) on the model + SGD on the centre -- with no host synchronisation inside the
step and, under data parallelism, one NCCL all-reduce of the flat gradient buffer.
"""
import torch

from . import ops, parallel
from .feature_extraction import LFCC

BF16 = torch.bfloat16


class Trainer:
    def __init__(self, arch="resnet", enc_dim=256, feat_len=750, padding="repeat", lr=5e-4, beta_1=0.9,
                 beta_2=0.999, eps=1e-8, weight_decay=5e-4, r_real=0.9, r_fake=0.2, alpha=20.0,
                 weight_loss=1.0, device="cuda", process_group=None, seed=None):
        self.arch, self.feat_len, self.padding = arch, feat_len, padding
        self.lr, self.betas, self.eps, self.wd = lr, (beta_1, beta_2), eps, weight_decay
        self.r_real, self.r_fake, self.alpha, self.weight_loss = r_real, r_fake, alpha, weight_loss
        self.device = torch.device(device)
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.lfcc = LFCC(320, 160, 512, 16000, 20).to(self.device)
        if arch == "resnet":
            from .engine import ResNetEngine
            self.engine = ResNetEngine(enc_dim=enc_dim, nclasses=2, device=self.device)
            self.layout = "resnet"
        elif arch == "ecapa":
            from .engine_ecapa import EcapaEngine
            self.engine = EcapaEngine(device=self.device)
            self.layout = "ecapa"
        else:
            raise ValueError(arch)
        if seed is not None:
            self.engine.init_parameters(seed)
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        c = torch.randn(1, enc_dim, generator=g)
        torch.nn.init.kaiming_uniform_(c, 0.25, generator=g)          # loss.py:183-184
        self.center = c.to(self.device)
        self.center_grad = torch.zeros_like(self.center)
        self.loss = torch.zeros(1, device=self.device)
        self.ce = torch.zeros(1, device=self.device)
        self.x0 = None
        self.dfeat = None
        self.score = None
        self.launches = 0
        self.adv, self.lr_d, self.adv_stats, self.adv_stats_c = [], 1e-4, [], []              # --ADV_AUG heads (attach_adversaries)
        self.reducer = None
        if self.world > 1:
            st = self.engine.store
            parallel.broadcast_state([st.params, self.center] + list(self.engine.buffers.named_f32().values()), 0, self.pg)
            self.engine.mark_dirty()
            self.reducer = parallel.GradReducer(st.grads, st.n_train, self.pg)
            self.engine.grad_hook = self.reducer.ready

    def load_state(self, model_sd, center=None):
        """Load reference-layout model weights (and the OC-Softmax centre)."""
        self.engine.load_state(model_sd)
        if center is not None:
            self.center.copy_(center.to(self.device))

    # ------------------------------------------------------------------------------------
    def features(self, waves, lengths=None, start=None):
        B = waves.shape[0]
        shape = (B, 1, 60, self.feat_len) if self.layout == "resnet" else (B, self.feat_len, 64)
        if self.x0 is None or tuple(self.x0.shape) != shape:
            self.x0 = torch.zeros(shape, device=self.device, dtype=BF16)
        # algorithmic bytes of the fused front-end: read the fp32 wave once, write the padded bf16 model-layout features once
        nbytes = float(waves.numel() * 4 + self.x0.numel() * 2)
        with ops.region("lfcc", nbytes=nbytes):
            self.lfcc.extract(waves, lengths=lengths, feat_len=self.feat_len, padding=self.padding, start=start,
                              layout=self.layout, dtype=BF16, out=self.x0)
        return self.x0[:, 0] if self.layout == "resnet" else self.x0

    def attach_adversaries(self, class_counts, lambda_=0.05, lr_d=1e-4, seed=None):
        """The channel classifier(s) of --ADV_AUG (main_train.py:211-224): one head per entry of `class_counts`
        (LA_aug / DF_aug: [n_codecs]; LAPA / DFPA: [n_codecs, n_devices]), Adam(lr_d, L2 5e-4) each."""
        from .adv import ChannelClassifier
        if seed is not None:
            torch.manual_seed(seed)
        self.adv = [ChannelClassifier(self.center.shape[1], c, lambda_, device=self.device) for c in class_counts]
        self.lr_d = lr_d
        return self.adv

    def train_step(self, waves, labels, lengths=None, start=None, lr=None, channels=None, step_seed=0):
        """One optimiser step on a (B, L) fp32 wave batch; returns the device loss tensor (no sync).
        channels: (B,) or (B, n_heads) int labels -- when given and adversaries are attached, the step follows
        main_train.py:377-453: the gradient-reversed channel loss joins the feature loss, then the encoder runs a
        second time and each classifier takes its own Adam step on the detached features."""
        eng = self.engine
        B = waves.shape[0]
        x0 = self.features(waves, lengths, start)
        feat, logits = eng.forward(x0, training=True)
        if self.dfeat is None or self.dfeat.shape[0] != B:
            self.dfeat = torch.empty(B, feat.shape[1], device=self.device)
            self.score = torch.empty(B, device=self.device)
        eng.zero_grad()
        self.center_grad.zero_()
        ops.ocsoftmax(feat, labels, self.center, B, feat.shape[1], self.r_real, self.r_fake, self.alpha,
                      self.weight_loss, self.loss, self.score, self.dfeat, self.center_grad, logits, logits.shape[1], self.ce)
        adv_on = channels is not None and len(self.adv) > 0
        if adv_on:
            ch = channels.view(B, -1)
            self.adv_stats = [clf.head_loss_and_feat_grad(feat, ch[:, i], self.dfeat, seed=2 * step_seed * len(self.adv) + i)
                              for i, clf in enumerate(self.adv)]
        if self.reducer is not None:
            self.reducer.begin()
        eng.backward(self.dfeat)
        scale = self.reducer.finish(self.center_grad) if self.reducer is not None else 1.0
        lr = self.lr if lr is None else lr
        eng.store.adam_step(lr, self.betas[0], self.betas[1], self.eps, self.wd, grad_scale=scale)
        ops.sgd_step(self.center, self.center_grad, self.center.numel(), lr, scale)
        if adv_on:                                                       # main_train.py:420-453
            feat2, _ = eng.forward(x0, training=True)
            self.adv_stats_c = [clf.classifier_step(feat2, ch[:, i], self.lr_d, self.betas[0], self.betas[1], self.eps, 0.0005,
                                                    seed=(2 * step_seed + 1) * len(self.adv) + i, group=self.pg)
                                for i, clf in enumerate(self.adv)]
        return self.loss

    @torch.no_grad()
    def eval_loss(self, waves, labels, lengths=None, start=None):
        """Validation pass (main_train.py:489-577): eval-mode forward, OC-Softmax loss and scores, no update."""
        eng = self.engine
        B = waves.shape[0]
        feat, logits = eng.forward(self.features(waves, lengths, start), training=False)
        loss = torch.empty(1, device=self.device)
        score = torch.empty(B, device=self.device)
        ops.ocsoftmax(feat, labels, self.center, B, feat.shape[1], self.r_real, self.r_fake, self.alpha, 1.0,
                      loss, score, None, None)
        return loss, score

    # ---- checkpoints (main_train.py:674-706 saves whole-module pickles) -------------------------
    def modules(self):
        """(feat_model, loss_model): the drop-in nn.Modules bound to this trainer's parameters."""
        from .loss import AngularIsoLoss
        if self.arch == "resnet":
            from .resnet import ResNet
            model = ResNet(3, self.engine.enc_dim, '18', nclasses=2, engine=self.engine)
        else:
            from .ecapa_tdnn import Res2Net2, Bottle2neck
            model = Res2Net2(Bottle2neck, C=self.engine.C, model_scale=self.engine.scale, nOut=self.engine.n_out,
                             n_mels=self.engine.n_mels, engine=self.engine)
        loss_model = AngularIsoLoss(self.center.shape[1], r_real=self.r_real, r_fake=self.r_fake, alpha=self.alpha)
        loss_model.center.data = self.center
        return model, loss_model

    def load_modules(self, model, loss_model=None):
        """Adopt the weights of (unpickled) drop-in modules, e.g. for --continue_training / scoring."""
        self.engine.load_state({k: v for k, v in model.state_dict().items()})
        if loss_model is not None:
            self.center.copy_(loss_model.center.detach().to(self.device))

    @torch.no_grad()
    def score_step(self, waves, lengths=None, start=None):
        """generate_score.py:84-119: eval-mode forward, returns +cos(feat, centre) per utterance."""
        eng = self.engine
        B = waves.shape[0]
        x0 = self.features(waves, lengths, start)
        feat, logits = eng.forward(x0, training=False)
        if self.score is None or self.score.shape[0] != B:
            self.score = torch.empty(B, device=self.device)
        ops.ocsoftmax(feat, None, self.center, B, feat.shape[1], self.r_real, self.r_fake, self.alpha, 1.0,
                      None, self.score, None, None)
        return -self.score

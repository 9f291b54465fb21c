"""Drop-in module: same constructor, parameter names / state_dict keys and `mc_inference`
signature as the reference `MultiHeadGatedAttentionMIL` (/root/reference/model.py:134-328).

Only the head is B200-native: the ResNet feature extractor runs once per bag in PyTorch
(north_star; /root/reference/model.py:276-277) and its (N,512) output goes straight into the
fused CUDA head — T MC-dropout passes for `mc_inference`, one all-keep pass for the eval-mode
`forward` (/root/reference/model.py:211-253).  `forward` in training mode keeps the plain
torch graph (autograd) so existing training scripts still run.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .head import HeadWeights, MCHeadResult, aux_pairwise_loss, head_forward_eval, mc_head


class Identity(nn.Module):
    def forward(self, x):
        return x


class AuxiliaryLoss(nn.Module):
    """Same surface as the reference's AuxiliaryLoss (/root/reference/model.py:405-438): `loss_type` 'pairwise'
    (L2 distance between the positive and the negative head's attention, hinged at `margin` for positive bags) or
    'cosine'; callers read `.scale` (model.py:246,322).  Plain torch: it is the training-time regulariser; the
    inference paths compute the pairwise form on the device with mcmil_aux_pairwise_loss."""

    def __init__(self, loss_type="pairwise", margin=1.0, scale=1.0):
        super().__init__()
        self.loss_type, self.margin, self.scale = loss_type, margin, scale

    def forward(self, pos_attention, neg_attention, is_positive):
        if self.loss_type == "pairwise":
            d = F.pairwise_distance(pos_attention, neg_attention, p=2)
            return torch.mean((self.margin - d).clamp(min=0)) if is_positive else torch.mean(d)
        if self.loss_type == "cosine":
            cs = F.cosine_similarity(pos_attention, neg_attention, dim=1)
            return torch.mean(cs) if is_positive else torch.mean(1 - cs)
        raise ValueError(f"Unknown loss type: {self.loss_type}")


def _make_backbone(backbone: str, pretrained: bool):
    import torchvision.models as tvm
    ctor = {"r18": tvm.resnet18, "r34": tvm.resnet34, "r50": tvm.resnet50}
    if not pretrained:
        return tvm.resnet18()                       # the reference ignores `backbone` here (model.py:176-177)
    weights = {"r18": tvm.ResNet18_Weights.IMAGENET1K_V1, "r34": tvm.ResNet34_Weights.IMAGENET1K_V1,
               "r50": tvm.ResNet50_Weights.IMAGENET1K_V1}[backbone]
    return ctor[backbone](weights=weights)


class _ExtractorRunner:
    """SURVEY.md §8f-4.  The ResNet extractor stays in PyTorch (north_star); this runner only changes HOW torch
    runs it: channels-last activations / weights and, in "graph" mode, one CUDA graph per bag shape, so a bag
    costs one graph launch instead of ~120 kernel launches.  BatchNorm keeps the reference's whole-bag batch
    statistics (`deactivate_batchnorm`: the bag is ONE batch inside the graph), results are the eager ones up
    to cuDNN's algorithm choice."""

    def __init__(self):
        self.graphs = {}          # (shape, device) -> (graph, static_in, static_out)
        self.fingerprint = None   # storage of the extractor's parameters / buffers the graphs were captured against

    @staticmethod
    def _fingerprint(module: nn.Module):
        """A captured graph bakes in the addresses of the parameters and buffers: any move / cast / re-assignment
        of them (model.cpu().cuda(), .half(), load_state_dict(assign=True)) must drop the graphs."""
        return tuple((t.data_ptr(), t.dtype, tuple(t.shape)) for t in list(module.parameters()) + list(module.buffers()))

    def run(self, module: nn.Module, x: torch.Tensor, mode: str) -> torch.Tensor:
        if mode == "eager":
            return module(x)
        if mode not in ("channels_last", "graph"):
            raise ValueError(f"extractor_mode must be eager / channels_last / graph, got {mode!r}")
        if x.device.type != "cuda":
            raise RuntimeError("extractor_mode channels_last / graph need CUDA tensors")
        if any(p.dim() == 4 and not p.is_contiguous(memory_format=torch.channels_last) for p in module.parameters()):
            module.to(memory_format=torch.channels_last)      # (re-done after any move that reset the layout)
        x = x.contiguous(memory_format=torch.channels_last)
        if mode == "channels_last":
            return module(x)
        fp = self._fingerprint(module)
        if fp != self.fingerprint:
            self.graphs.clear()
            self.fingerprint = fp
        key = (tuple(x.shape), str(x.device))
        ent = self.graphs.get(key)
        if ent is None:
            if len(self.graphs) >= 8:
                self.graphs.clear()
            static_in = x.clone(memory_format=torch.channels_last)
            side = torch.cuda.Stream(x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(2):
                    module(static_in)                       # warm-up: cuDNN algorithm selection, workspaces
            torch.cuda.current_stream(x.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                static_out = module(static_in)
            ent = self.graphs[key] = (g, static_in, static_out)
        g, static_in, static_out = ent
        static_in.copy_(x)
        g.replay()
        return static_out.clone()


class MultiHeadGatedAttentionMIL(nn.Module):
    def __init__(self, num_classes=2, backbone="r18", pretrained=True, L=512, D=128,
                 feature_dropout=0.1, attention_dropout=0.1, shared_attention=True, neptune_run=None):
        super().__init__()
        if L != 512 or D != 128:
            raise NotImplementedError("the B200 head is built for L=512, D=128 (the reference defaults)")
        self.auxiliary_loss = AuxiliaryLoss(loss_type="pairwise", margin=1.0, scale=.5)      # model.py:149-151
        self.fold_idx = None
        self.neptune_run = neptune_run
        self.L, self.D = L, D
        self.num_classes = num_classes
        self.shared_attention = shared_attention
        self.feature_extractor = _make_backbone(backbone, pretrained)
        self.feature_extractor.fc = Identity()
        if shared_attention:
            self.attention_V = nn.Sequential(nn.Linear(L, D), nn.Tanh())
            self.attention_U = nn.Sequential(nn.Linear(L, D), nn.Sigmoid())
        else:
            self.attention_V = nn.ModuleList([nn.Sequential(nn.Linear(L, D), nn.Tanh()) for _ in range(num_classes)])
            self.attention_U = nn.ModuleList([nn.Sequential(nn.Linear(L, D), nn.Sigmoid()) for _ in range(num_classes)])
        self.attention_weights = nn.ModuleList([nn.Linear(D, 1) for _ in range(num_classes)])
        self.classifiers = nn.ModuleList([nn.Linear(L, 1, bias=False) for _ in range(num_classes)])
        self.feature_dropout = nn.Dropout(feature_dropout)
        self.attention_dropouts = nn.ModuleList([nn.Dropout(attention_dropout) for _ in range(num_classes)])
        self._packed = None          # (HeadWeights, version key)
        self.mc_seed = 0             # Philox key of the next mc_inference call (auto-incremented)
        self.last_result: MCHeadResult | None = None
        # Eval-mode, no-grad, CUDA forward() goes through the fused head (forward_eval_fused): features and weights
        # are rounded to fp16 for the tensor cores and tanh is the hardware approximation, so logits differ from the
        # fp32 torch graph by up to ~2e-3 absolute (attention ~1e-4 relative).  Set False to keep the plain torch
        # graph, or `fused_eval_impl = "simt_fp32"` for the fused path in full fp32.
        self.fused_eval = True
        self.fused_eval_impl = "tcgen05"
        self.validate_features = False   # True: raise if the extractor's features are not finite or exceed fp16 range
        self.extractor_mode = "eager"  # "channels_last" / "graph": how torch runs the extractor for mc_inference
        self._extractor_runner = _ExtractorRunner()

    # ------------------------------------------------------------------ forward (model.py:211-253)
    @property
    def AUX_MARGIN(self):
        return float(self.auxiliary_loss.margin)

    @property
    def AUX_SCALE(self):
        return float(self.auxiliary_loss.scale)

    def _aux_fused_ok(self):
        return self.auxiliary_loss.loss_type == "pairwise"

    def forward(self, x, targets=None):
        """Training keeps the plain-torch graph (autograd).  Deterministic inference — eval mode, grad
        disabled, CUDA input, as in the reference's validate / test loops (net_utils.py:82-114,160-192) —
        runs the head through the fused kernels (`forward_eval_fused`)."""
        if self.fused_eval and not self.training and not torch.is_grad_enabled() and x.is_cuda:
            return self.forward_eval_fused(x, targets)
        bs, n, ch, w, h = x.shape
        H = self.feature_extractor(x.view(bs * n, ch, w, h))
        H = self.feature_dropout(H).view(bs, n, -1)
        M, A_all = [], []
        for i in range(self.num_classes):
            V = self.attention_V if self.shared_attention else self.attention_V[i]
            U = self.attention_U if self.shared_attention else self.attention_U[i]
            a = self.attention_weights[i](V(H) * U(H)).transpose(2, 1)
            a = F.softmax(self.attention_dropouts[i](a), dim=2)
            A_all.append(a)
            M.append(torch.matmul(a, H))
        M = torch.cat(M, dim=1)
        A_all = torch.cat(A_all, dim=1)
        Y = torch.cat([self.classifiers[i](M[:, i, :]) for i in range(self.num_classes)], dim=-1)
        aux = None
        if targets is not None:                     # model.py:243-248
            aux = self.auxiliary_loss.scale * self.auxiliary_loss(A_all[:, 1, :], A_all[:, 0, :], targets.item() == 1)
        return Y, A_all, aux

    def forward_eval_fused(self, x, targets=None):
        """model.py:211-253 in eval mode (dropout inactive): extractor in torch, then ONE pass of the fused
        head with all-keep masks; the auxiliary loss (model.py:243-248) by mcmil_aux_pairwise_loss.
        Returns (Y (bs, C), A_all (bs, C, n), auxiliary_loss or None) like the reference."""
        bs, n = x.shape[:2]
        if x.device.type != "cuda":
            raise RuntimeError("forward_eval_fused needs CUDA tensors (there is no CPU fallback)")
        with torch.no_grad():
            H = self.feature_extractor(x.view(bs * n, *x.shape[2:])).view(bs * n, -1).float().contiguous()
            cu = [i * n for i in range(bs + 1)]
            Y, A = head_forward_eval(self._head_weights(x.device), H, cu, impl=self.fused_eval_impl,
                                     validate=self.validate_features)             # (bs, C), (C, bs*n)
            A_all = A.view(self.num_classes, bs, n).permute(1, 0, 2).contiguous()   # (bs, C, n)
            aux = None
            if targets is not None:
                if self.num_classes < 2:
                    raise RuntimeError("the auxiliary loss compares heads 1 and 0 (model.py:246-247)")
                if self._aux_fused_ok():
                    aux = aux_pairwise_loss(A.view(1, self.num_classes, bs * n), targets.item() == 1, cu,
                                            margin=self.AUX_MARGIN, scale=self.AUX_SCALE).mean()   # torch.mean over bs
                else:                                   # 'cosine' (model.py:428-438): tiny, torch on the device
                    aux = self.auxiliary_loss.scale * self.auxiliary_loss(A_all[:, 1, :], A_all[:, 0, :],
                                                                          targets.item() == 1)
        return Y, A_all, aux

    # ------------------------------------------------------------------ the B200 hot path
    def _head_weights(self, device) -> HeadWeights:
        params = [p for n, p in self.named_parameters() if not n.startswith("feature_extractor")]
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or self._packed[1] != key:
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith("feature_extractor")}
            self._packed = (HeadWeights(sd, device), key)
        return self._packed[0]

    def extract_features(self, input_tensor):
        """(1,N,3,H,W) -> (N,512): the extractor runs ONCE per bag (model.py:276-277)."""
        bs, n = input_tensor.shape[:2]
        if bs != 1:
            raise RuntimeError("mc_inference supports bs == 1 only (as the reference, model.py:309)")
        x = input_tensor.view(-1, *input_tensor.shape[-3:])
        mode = self.extractor_mode if (x.is_cuda and x.dim() == 4 and not torch.is_grad_enabled()) else "eager"
        H = self._extractor_runner.run(self.feature_extractor, x, mode)
        return H.view(n, -1).float().contiguous()

    def mc_inference_stats(self, input_tensor, N=30, device="cuda", seed=None, return_attention=False,
                           keep_f_bits=None, keep_a_bits=None, impl="tcgen05") -> MCHeadResult:
        """Features once, then N fused MC-dropout passes; returns logits + Welford statistics."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("the fused MC-dropout head needs a CUDA device (there is no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.eval()                                   # model.py:263
        self.to(device)                               # model.py:264
        for m in self.modules():                      # model.py:268-271: dropout stays in train mode
            if isinstance(m, nn.Dropout):
                m.train()
        if seed is None:
            seed, self.mc_seed = self.mc_seed, self.mc_seed + 1
        with torch.no_grad():
            H = self.extract_features(input_tensor.to(device))
            res = mc_head(self._head_weights(device), H, int(N), seed=seed,
                          p_f=float(self.feature_dropout.p), p_a=float(self.attention_dropouts[0].p),
                          keep_f_bits=keep_f_bits, keep_a_bits=keep_a_bits,
                          return_attention=return_attention, impl=impl, validate=self.validate_features)
        self.last_result = res
        return res

    def mc_inference(self, input_tensor, N=30, device="cuda", targets=None, seed=None, legacy_tuple=False):
        """Same contract as the reference (model.py:256-328): returns
        (Y (N,1,C) logits, A (N,1,C,num_instances)); `legacy_tuple=True` appends the third
        value (None) that the reference's committed callers unpack (infer.py:191,
        net_utils.py:126,205): the list of N per-pass auxiliary losses when `targets` is given
        (model.py:318-326; computed by mcmil_aux_pairwise_loss), else None.  The Welford
        statistics of the same call are kept in `self.last_result`."""
        res = self.mc_inference_stats(input_tensor, N=N, device=device, seed=seed, return_attention=True)
        Y = res.Y.permute(1, 0, 2).contiguous()                     # (N, 1, C)
        A = res.A.unsqueeze(1)                                      # (N, 1, C, n)
        if not legacy_tuple:
            return Y, A
        losses = None
        if targets is not None and self.num_classes >= 2:           # model.py:318-326: one loss per MC pass
            if self._aux_fused_ok():
                per_pass = aux_pairwise_loss(res.A, targets.item() == 1, margin=self.AUX_MARGIN, scale=self.AUX_SCALE)[0]
                losses = list(per_pass.unbind(0))
            else:
                losses = [self.auxiliary_loss.scale * self.auxiliary_loss(res.A[i, 1:2], res.A[i, 0:1], targets.item() == 1)
                          for i in range(res.A.shape[0])]
        return Y, A, losses

    def mc_inference_serial(self, input_tensor, N=30, device="cuda"):
        """The reference's second MC entry point (/root/reference/model.py:330-401): N bag-sized passes in a Python
        loop (~20 s per bag on the reference's CPU path).  Same distribution of (predictions (N,bs,C),
        attention_weights (N,bs,C,n)) as `mc_inference`; here it is the same fused batched call (the reference's
        two paths also draw different samples for the same seed: they consume the RNG in different orders)."""
        return self.mc_inference(input_tensor, N=N, device=device)


def deactivate_batchnorm(m):
    """What every reference script applies before loading weights (main.py:16-20,
    infer.py:105-109): BatchNorm uses per-bag batch statistics even in eval."""
    if isinstance(m, nn.BatchNorm2d):
        m.track_running_stats = False
        m.running_mean = None
        m.running_var = None

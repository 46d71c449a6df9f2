"""Plain PyTorch fp32 reference of the forward stage (TEST INFRASTRUCTURE ONLY).

Mirrors the reference's model call contract exactly (src/inference.rs:147-175):

    quals  = u8 -> f32 ; quals * (2/93) - (66/93 + 1)          (:19-21,152-153, two fp32 ops)
    bases  = u8 -> i32                                          (:156)
    (info_logits [sum lens], bases_logits [sum lens, 5]) =
        model(bases [B,L,31], quals [B,L,31], lens [B] i32, indices List[i32 tensor])   (:155-172)

The graph itself is NOT in the reference repository (it ships inside an external
TorchScript archive); the architecture here is the one BASELINE.json's
north_star names (conv stem, read-axis MHA, per-position FFN, base + info heads) with the
tensor names/shapes of herro_b200/weights.py.  Like the reference's TorchScript graph it
evaluates the stem over the whole [B, L, 31] batch tensor (including the batch padding rows,
token 11 / qual byte 126, src/inference.rs:86-97) and only then gathers `indices`.

`HerroNet` is scriptable (torch.jit.script) so tools/make_torchscript.py can emit an archive
the unmodified reference binary would accept through `CModule::forward_is`.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

QUAL_SCALE = 2.0 / 93.0             # src/inference.rs:19-20
QUAL_OFFSET = 2.0 * 33.0 / 93.0 + 1.0  # src/inference.rs:21


class EncoderLayer(nn.Module):
    """Pre-LN transformer encoder layer over the read axis (S = 31 tokens per position)."""

    def __init__(self, C: int, H: int, Fd: int):
        super().__init__()
        self.C, self.H = C, H
        self.ln1 = nn.LayerNorm(C, eps=1e-5)
        self.qkv = nn.Linear(C, 3 * C)
        self.out = nn.Linear(C, C)
        self.ln2 = nn.LayerNorm(C, eps=1e-5)
        self.ff1 = nn.Linear(C, Fd)
        self.ff2 = nn.Linear(Fd, C)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # [N, S, C]
        N, S, C = x.shape
        H = self.H
        dh = C // H
        h = self.ln1(x)
        qkv = self.qkv(h).view(N, S, 3, H, dh)
        q = qkv[:, :, 0].transpose(1, 2)  # [N,H,S,dh]
        k = qkv[:, :, 1].transpose(1, 2)
        v = qkv[:, :, 2].transpose(1, 2)
        att = torch.matmul(q, k.transpose(-1, -2)) * (1.0 / math.sqrt(dh))
        att = torch.softmax(att, dim=-1)
        o = torch.matmul(att, v).transpose(1, 2).reshape(N, S, C)
        x = x + self.out(o)
        h = self.ln2(x)
        x = x + self.ff2(F.relu(self.ff1(h)))
        return x


class HerroNet(nn.Module):
    def __init__(self, stem_k: int = 33, channels: int = 128, heads: int = 4, layers: int = 2, ffn: int = 512,
                 collapse: int = 256):
        super().__init__()
        self.R = 31
        self.embedding = nn.Embedding(12, 6, padding_idx=11)
        self.stem = nn.Conv2d(7, channels, (stem_k, 1), padding=(stem_k // 2, 0))
        self.read_pos = nn.Parameter(torch.zeros(31, channels))
        self.layers = nn.ModuleList([EncoderLayer(channels, heads, ffn) for _ in range(layers)])
        self.lnf = nn.LayerNorm(channels, eps=1e-5)
        self.collapse = nn.Linear(31 * channels, collapse)
        self.base_head = nn.Linear(collapse, 5)
        self.info_head = nn.Linear(collapse, 1)

    def stem_features(self, bases: torch.Tensor, quals: torch.Tensor) -> torch.Tensor:
        x = self.embedding(bases)                               # [B,L,31,6]
        x = torch.cat([x, quals.unsqueeze(-1)], dim=-1)         # [B,L,31,7]
        x = x.permute(0, 3, 1, 2)                               # [B,7,L,31]
        x = F.relu(self.stem(x))                                # [B,C,L,31]
        return x.permute(0, 2, 3, 1)                            # [B,L,31,C]

    def forward(self, bases: torch.Tensor, quals: torch.Tensor, lens: torch.Tensor,
                indices: List[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        x = self.stem_features(bases, quals)
        sel: List[torch.Tensor] = []
        for b in range(len(indices)):
            sel.append(x[b].index_select(0, indices[b].to(torch.long)))
        t = torch.cat(sel, dim=0) + self.read_pos               # [N,31,C]
        for layer in self.layers:
            t = layer(t)
        t = self.lnf(t)
        z = F.relu(self.collapse(t.reshape(t.shape[0], -1)))
        return self.info_head(z).squeeze(-1), self.base_head(z)


def from_weights(cfg, tensors: dict) -> HerroNet:
    """Build the module from a herro_b200.weights blob (cfg: NetConfig, tensors: name->ndarray)."""
    net = HerroNet(cfg.stem_k, cfg.channels, cfg.heads, cfg.layers, cfg.ffn, cfg.collapse)
    T = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in tensors.items()}
    with torch.no_grad():
        net.embedding.weight.copy_(T["emb"])
        net.stem.weight.copy_(T["stem_w"].unsqueeze(-1))        # [C,7,K] -> [C,7,K,1]
        net.stem.bias.copy_(T["stem_b"])
        net.read_pos.copy_(T["read_pos"])
        for l, layer in enumerate(net.layers):
            p = f"l{l}."
            layer.ln1.weight.copy_(T[p + "ln1_g"]); layer.ln1.bias.copy_(T[p + "ln1_b"])
            layer.qkv.weight.copy_(T[p + "wqkv"]); layer.qkv.bias.copy_(T[p + "bqkv"])
            layer.out.weight.copy_(T[p + "wo"]); layer.out.bias.copy_(T[p + "bo"])
            layer.ln2.weight.copy_(T[p + "ln2_g"]); layer.ln2.bias.copy_(T[p + "ln2_b"])
            layer.ff1.weight.copy_(T[p + "w1"]); layer.ff1.bias.copy_(T[p + "b1"])
            layer.ff2.weight.copy_(T[p + "w2"]); layer.ff2.bias.copy_(T[p + "b2"])
        net.lnf.weight.copy_(T["lnf_g"]); net.lnf.bias.copy_(T["lnf_b"])
        net.collapse.weight.copy_(T["wc"]); net.collapse.bias.copy_(T["bc"])
        net.base_head.weight.copy_(T["wb"]); net.base_head.bias.copy_(T["bb"])
        net.info_head.weight.copy_(T["wi"]); net.info_head.bias.copy_(T["bi"])
    net.eval()
    return net


@torch.no_grad()
def run_batch(net: HerroNet, bases_u8: np.ndarray, quals_u8: np.ndarray, lens: np.ndarray, indices: list):
    """The reference's `inference()` (src/inference.rs:147-175) on CPU, fp32."""
    quals = torch.from_numpy(quals_u8).to(torch.float32)
    quals = QUAL_SCALE * quals - QUAL_OFFSET                    # mul then sub, fp32 (H11)
    bases = torch.from_numpy(bases_u8).to(torch.int32)
    lens_t = torch.from_numpy(np.asarray(lens, dtype=np.int32))
    idx = [torch.from_numpy(np.asarray(i, dtype=np.int32)) for i in indices]
    info, bl = net(bases, quals, lens_t, idx)
    sizes = [int(l) for l in lens]
    return ([t.numpy() for t in torch.split(info, sizes)], [t.numpy() for t in torch.split(bl, sizes)])

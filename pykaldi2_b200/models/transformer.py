"""Transformer acoustic model (BASELINE config 5).

Mirrors the reference's ``models.transformer.TransformerAM`` (models/transformer.py:52-94): same constructor
signature, same ``forward(data[T,B,F], src_mask, src_key_padding_mask) -> [T,B,N]`` convention and the same
state-dict keys (``input_layer``, ``transformer.layers.<i>.encoder_layer.{self_attn,linear1,linear2,norm1,norm2}``,
``transformer.layers.<i>.conv1d``, ``transformer.norm``, ``output_layer``, buffer ``pos_encoder.pe``), so
checkpoints move both ways.  Architecture per layer (reference :52-68): post-norm encoder layer (self-attention,
ReLU feed-forward, LayerNorm eps 1e-5) followed by Conv1d(k, stride, padding=1) over time and ReLU; a final
LayerNorm; the positional encoding is constructed but NOT applied (commented out at :88 of the reference).

Not a port: the layers are written out on ``F.scaled_dot_product_attention`` in a batch-first layout (one
additive mask built once per forward from the look-ahead mask and the key-padding mask) and run under bf16
autocast on the GPU -- SURVEY 8f-3 scopes this model to stock torch kernels; the hand-written kernels of this
repo are on the BLSTM path.  The sequence losses (``ops.MMIFunction`` / ``ops.sMBRFunction``) are shared.
"""
import math

import torch as th
import torch.nn as nn
import torch.nn.functional as F


class PositionalEncoding(nn.Module):
    """Sinusoidal table kept as the ``pe`` buffer ([max_len, 1, dim]) for checkpoint compatibility."""

    def __init__(self, dim_model, dropout=0, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pos = th.arange(max_len, dtype=th.float32)[:, None]
        freq = th.exp(th.arange(0, dim_model, 2, dtype=th.float32) * (-math.log(10000.0) / dim_model))
        pe = th.zeros(max_len, 1, dim_model)
        pe[:, 0, 0::2] = th.sin(pos * freq)
        pe[:, 0, 1::2] = th.cos(pos * freq)
        self.register_buffer("pe", pe)

    def forward(self, x):                      # x: [T, B, D]
        return self.dropout(x + self.pe[:x.size(0)])


class _SelfAttention(nn.Module):
    """Parameters named as nn.MultiheadAttention's (in_proj_weight / in_proj_bias / out_proj)."""

    def __init__(self, dim_model, nheads, dropout):
        super().__init__()
        if dim_model % nheads:
            raise ValueError("dim_model must be divisible by nheads")
        self.nheads, self.p = nheads, dropout
        self.in_proj_weight = nn.Parameter(th.empty(3 * dim_model, dim_model))
        self.in_proj_bias = nn.Parameter(th.zeros(3 * dim_model))
        self.out_proj = nn.Linear(dim_model, dim_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    def forward(self, x, mask):                # x: [B, T, D]; mask: additive [B, 1, T, T] or None
        B, T, D = x.shape
        h = self.nheads
        qkv = F.linear(x, self.in_proj_weight, self.in_proj_bias).view(B, T, 3, h, D // h).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=mask,
                                           dropout_p=self.p if self.training else 0.0)
        return self.out_proj(o.transpose(1, 2).reshape(B, T, D))


class _EncoderLayer(nn.Module):
    """Post-norm encoder layer, ReLU feed-forward (what nn.TransformerEncoderLayer defaults to)."""

    def __init__(self, dim_model, nheads, dim_feedforward, dropout):
        super().__init__()
        self.self_attn = _SelfAttention(dim_model, nheads, dropout)
        self.linear1 = nn.Linear(dim_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, dim_model)
        self.norm1 = nn.LayerNorm(dim_model)
        self.norm2 = nn.LayerNorm(dim_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)

    def forward(self, x, mask):
        x = self.norm1(x + self.dropout1(self.self_attn(x, mask)))
        return self.norm2(x + self.dropout2(self.linear2(self.dropout(F.relu(self.linear1(x))))))


class TransformerEncoderLayerWithConv1d(nn.Module):
    def __init__(self, dim_model, nheads, dim_feedforward, dropout, kernel_size, stride):
        super().__init__()
        self.encoder_layer = _EncoderLayer(dim_model, nheads, dim_feedforward, dropout)
        self.conv1d = nn.Conv1d(dim_model, dim_model, kernel_size, stride=stride, padding=1)

    def forward(self, x, mask):                # [B, T, D]
        x = self.encoder_layer(x, mask)
        return F.relu(self.conv1d(x.transpose(1, 2))).transpose(1, 2)


class _Encoder(nn.Module):
    """``layers`` + final ``norm`` (the attribute names of nn.TransformerEncoder)."""

    def __init__(self, make_layer, nlayers, dim_model):
        super().__init__()
        self.layers = nn.ModuleList([make_layer() for _ in range(nlayers)])
        # nn.TransformerEncoder deep-copies ONE layer: every layer of the reference starts from the same weights
        for layer in self.layers[1:]:
            layer.load_state_dict(self.layers[0].state_dict())
        self.norm = nn.LayerNorm(dim_model)


class TransformerAM(nn.Module):
    def __init__(self, dim_feat, dim_model, nheads, dim_feedforward, nlayers, dropout, output_size,
                 kernel_size=3, stride=1):
        super().__init__()
        if stride != 1 or kernel_size != 3:
            # the reference's masks assume the frame count is unchanged (padding = 1 only does that for k = 3, stride = 1)
            raise ValueError("TransformerAM: only kernel_size=3, stride=1 keep the frame count (reference default)")
        self.pos_encoder = PositionalEncoding(dim_model, dropout)
        self.input_layer = nn.Linear(dim_feat, dim_model)
        self.output_layer = nn.Linear(dim_model, output_size)
        self.transformer = _Encoder(
            lambda: TransformerEncoderLayerWithConv1d(dim_model, nheads, dim_feedforward, dropout, kernel_size, stride),
            nlayers, dim_model)
        self.autocast_bf16 = True              # bf16 operands on the GPU; parameters and the output stay fp32

    @staticmethod
    def _additive_mask(T, src_mask, key_padding_mask, dtype, device):
        mask = None
        if src_mask is not None:               # [T, T]: float additive (0 / -inf) or bool (True = masked)
            m = src_mask.to(device)
            if m.dtype == th.bool:
                m = th.zeros(T, T, dtype=dtype, device=device).masked_fill(m, float("-inf"))
            mask = m.to(dtype)[None, None]
        if key_padding_mask is not None:       # [B, T] bool, True = padding
            k = th.zeros(key_padding_mask.shape, dtype=dtype, device=device)
            k = k.masked_fill(key_padding_mask.to(device), float("-inf"))[:, None, None, :]
            mask = k if mask is None else mask + k
        return mask

    def forward(self, data, src_mask=None, src_key_padding_mask=None):
        """data: [T, B, F] (the reference's time-major convention) -> logits [T, B, N] float32."""
        T = data.size(0)
        amp = self.autocast_bf16 and data.is_cuda
        with th.autocast("cuda", dtype=th.bfloat16, enabled=amp):
            x = self.input_layer(data.transpose(0, 1))                     # [B, T, D]
            mask = self._additive_mask(T, src_mask, src_key_padding_mask, x.dtype, x.device)
            if mask is not None and mask.size(0) == 1:
                mask = mask.expand(x.size(0), 1, T, T)
            for layer in self.transformer.layers:
                x = layer(x, mask)
            x = self.output_layer(self.transformer.norm(x))
        return x.float().transpose(0, 1)


def look_ahead_mask(T, look_ahead, device=None):
    """Additive [T, T] mask of bin/train_transformer_se.py:246-249: query t sees keys <= t + look_ahead."""
    keep = th.tril(th.ones(T, T, device=device), diagonal=look_ahead)
    return th.zeros(T, T, device=device).masked_fill(keep == 0, float("-inf"))

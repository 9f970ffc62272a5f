"""Mirror of the reference's transformer.py on the B200 kernels.

Same class names, constructor signatures, parameter names and shapes as the reference
(TransformerEncoderLayer transformer.py:7-60, MultiHeadAttention :62-112,
LearnedRelativePositionalEmbedding :114-305) so that checkpoints interchange; the arithmetic is
the closed form of SURVEY.md §8 a4-a6 executed by csrc/{gemm_simt,attn,norm}.cu:
  fused QKV GEMM -> positional-logit GEMM per head -> band attention kernel -> out-proj GEMM ->
  add+dropout+LayerNorm kernel -> FFN (GEMM+bias+ReLU+dropout, GEMM+bias) -> add+dropout+LN.
Internally activations are token-major (B*T, D); `forward(src)` keeps the reference's
seq-first (T, B, D) contract for stand-alone use.
"""
import copy
import random

import torch
from torch import nn

from . import functional as F_


def _fresh_seed():
    """64-bit Philox key for one forward pass, drawn from torch's CPU generator so that
    torch.manual_seed() controls dropout like it does in the reference."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


class LearnedRelativePositionalEmbedding(nn.Module):
    """Holds the (num_heads, 2*max_relative_pos-1, embedding_dim, 1) table of
    transformer.py:146-160.  The table never receives a gradient in the reference (it is
    re-padded under no_grad, transformer.py:213-219 — SURVEY.md F3); the same holds here
    because the band-attention op treats it as a constant."""

    def __init__(self, max_relative_pos, num_heads, embedding_dim, unmasked=False,
                 heads_share_embeddings=False, add_to_values=False):
        super().__init__()
        if not unmasked or heads_share_embeddings or add_to_values:
            raise NotImplementedError("only the encoder configuration the reference instantiates "
                                      "(transformer.py:83) is built")
        self.max_relative_pos = max_relative_pos
        self.num_heads = num_heads
        self.embedding_dim = embedding_dim
        self.unmasked = unmasked
        self.heads_share_embeddings = heads_share_embeddings
        self.add_to_values = add_to_values
        num_embeddings = 2 * max_relative_pos - 1
        self.embeddings = nn.Parameter(torch.zeros(num_heads, num_embeddings, embedding_dim, 1))
        nn.init.normal_(self.embeddings, mean=0.0, std=embedding_dim ** (-0.5))

    def padded_table(self):
        """(H, RW, dh) constant with RW = 2*max_relative_pos rounded up to a multiple of 4.
        The table never trains (F3), so the padded copy - and the operand planes the attention
        kernels derive from it, which functional caches ON this tensor - are rebuilt only when
        the parameter is written (load_state_dict, .to())."""
        emb = self.embeddings
        key = (emb._version, emb.data_ptr(), emb.device)
        if getattr(self, "_table_key", None) != key:
            with torch.no_grad():
                e = emb[..., 0]
                rw = (e.shape[1] + 1 + 3) // 4 * 4
                table = torch.nn.functional.pad(e, (0, 0, 0, rw - e.shape[1])).contiguous()
            if torch.cuda.is_current_stream_capturing():
                return table            # graph-pool memory must not outlive the capture
            self._table, self._table_key = table, key
        return self._table


class MultiHeadAttention(nn.Module):
    def __init__(self, d_model=256, n_head=4, dropout=0.1, relative_positional=True,
                 relative_positional_distance=100):
        super().__init__()
        self.d_model = d_model
        self.n_head = n_head
        d_qkv = d_model // n_head
        assert d_qkv * n_head == d_model, 'd_model must be divisible by n_head'
        self.d_qkv = d_qkv
        self.w_q = nn.Parameter(torch.empty(n_head, d_model, d_qkv))
        self.w_k = nn.Parameter(torch.empty(n_head, d_model, d_qkv))
        self.w_v = nn.Parameter(torch.empty(n_head, d_model, d_qkv))
        self.w_o = nn.Parameter(torch.empty(n_head, d_qkv, d_model))
        for w in (self.w_q, self.w_k, self.w_v, self.w_o):   # transformer.py:75-78
            nn.init.xavier_normal_(w)
        self.dropout = nn.Dropout(dropout)
        if not relative_positional:
            raise NotImplementedError("the reference always uses relative positions "
                                      "(architecture.py:53)")
        self.relative_positional = LearnedRelativePositionalEmbedding(
            relative_positional_distance, n_head, d_qkv, True)

    def qkv_weight(self):
        """(D, 3D) GEMM-layout weight: column h*dh + a of block j holds w_j[h, :, a]."""
        D = self.d_model
        blocks = [w.permute(1, 0, 2).reshape(D, D) for w in (self.w_q, self.w_k, self.w_v)]
        return torch.cat(blocks, dim=1)

    def forward_tokens(self, x2d, B, T, seed=0, site=0, wp=None):
        """x2d: (B*T, D) token-major -> attention output (B*T, D) (before residual/LN).
        wp: the model's WeightPlanes arena (fused QKV / out-projection operands)."""
        H, dh, D = self.n_head, self.d_qkv, self.d_model
        p = self.dropout.p if self.training else 0.0
        M = x2d.shape[0]
        qf = wp.get(self, "qkv_f") if wp is not None else None
        arena = (qf is not None and F_._tc_fwd_ok(M, 3 * D, D) and F_._tc_fwd_ok(M, D, 3 * D)
                 and F_._tc_wgrad_ok(M, 3 * D, D) and F_._tc_wgrad_ok(M, D, D))
        if arena:
            qkv = F_._QKVFn.apply(x2d, self.w_q, self.w_k, self.w_v, qf, wp.get(self, "qkv_b"))
        else:
            qkv = F_.linear(x2d, self.qkv_weight(), None)
        W = self.relative_positional.max_relative_pos - 1
        o = F_.band_attention(qkv, self.relative_positional.padded_table(), B, T, H, dh, W, p, seed,
                              site)
        if arena:
            return F_._OutProjFn.apply(o, self.w_o, wp.get(self.w_o, "f"), wp.get(self.w_o, "b"))
        return F_.linear(o, self.w_o.reshape(D, D), None)

    def forward(self, x):
        """x: (length, batch, d_model) as in transformer.py:87-112."""
        T, B, D = x.shape
        y = self.forward_tokens(x.transpose(0, 1).reshape(B * T, D).contiguous(), B, T,
                                _fresh_seed() if self.training else 0, 0)
        return y.view(B, T, D).transpose(0, 1)


class TransformerEncoderLayer(nn.Module):
    """Post-norm encoder layer, transformer.py:7-60 (mask arguments accepted and ignored)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, relative_positional=True,
                 relative_positional_distance=100):
        super().__init__()
        self.self_attn = MultiHeadAttention(d_model, nhead, dropout=dropout,
                                            relative_positional=relative_positional,
                                            relative_positional_distance=relative_positional_distance)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.activation = nn.ReLU()

    def forward_tokens(self, x2d, B, T, seed=0, site0=0, wp=None):
        """x2d: (B*T, D) -> (B*T, D).  Dropout sites site0 .. site0+3 (probs, attn-out, ffn, ffn-out).
        With the WeightPlanes arena and tensor-core eligible shapes each half of the layer is ONE
        autograd node (functional._AttnBlockFn / _FFNBlockFn); otherwise the single ops compose."""
        tr = self.training
        at = self.self_attn
        H, dh, D = at.n_head, at.d_qkv, at.d_model
        M = x2d.shape[0]
        W = at.relative_positional.max_relative_pos - 1
        p_attn = at.dropout.p if tr else 0.0
        p1 = self.dropout1.p if tr else 0.0
        qf = wp.get(at, "qkv_f") if wp is not None else None
        if qf is not None and F_.attn_block_ok(M, B, T, H, dh, W, D):
            x1 = F_._AttnBlockFn.apply(
                x2d, at.w_q, at.w_k, at.w_v, at.w_o, at.relative_positional.padded_table(),
                self.norm1.weight, self.norm1.bias, int(B), int(T), int(W), float(p_attn), float(p1),
                int(seed), int(site0), float(self.norm1.eps), qf, wp.get(at, "qkv_b"),
                wp.get(at.w_o, "f"), wp.get(at.w_o, "b"))
        else:
            a = at.forward_tokens(x2d, B, T, seed, site0, wp)
            x1 = F_.add_dropout_layernorm(x2d, a, self.norm1.weight, self.norm1.bias, p1, seed,
                                          site0 + 1, self.norm1.eps)
        p_ffn = self.dropout.p if tr else 0.0
        p2 = self.dropout2.p if tr else 0.0
        w1, w2 = self.linear1.weight, self.linear2.weight
        pl = ([wp.get(w1, "f"), wp.get(w1, "b"), wp.get(w2, "f"), wp.get(w2, "b")]
              if wp is not None else [None])
        if all(t is not None for t in pl) and F_.ffn_block_ok(M, D, w1.shape[0]):
            return F_._FFNBlockFn.apply(x1, w1, self.linear1.bias, w2, self.linear2.bias,
                                        self.norm2.weight, self.norm2.bias, float(p_ffn), float(p2),
                                        int(seed), int(site0 + 2), float(self.norm2.eps), *pl)
        f = F_.ffn_native(x1, w1, self.linear1.bias, w2, self.linear2.bias, p_ffn, seed, site0 + 2, wp)
        return F_.add_dropout_layernorm(x1, f, self.norm2.weight, self.norm2.bias, p2, seed,
                                        site0 + 3, self.norm2.eps)

    def forward(self, src, src_mask=None, src_key_padding_mask=None, is_causal=False):
        T, B, D = src.shape
        y = self.forward_tokens(src.transpose(0, 1).reshape(B * T, D).contiguous(), B, T,
                                _fresh_seed() if self.training else 0, 0)
        return y.view(B, T, D).transpose(0, 1)


class TransformerEncoder(nn.Module):
    """Stand-in for torch.nn.TransformerEncoder as the reference uses it (architecture.py:54):
    `num_layers` deep copies of one layer (identical initial weights), no final norm.  Parameter
    names are `layers.{i}.*`, matching the reference's checkpoints."""

    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        # optional callable (layer_index, layer_input): training.OverlappedAllReduce hangs its
        # "gradients of layers >= i are complete" hooks on the layer inputs through it
        self.layer_input_hook = None

    def forward_tokens(self, x2d, B, T, wp=None):
        seed = _fresh_seed() if self.training else 0
        for i, layer in enumerate(self.layers):
            if self.layer_input_hook is not None:
                self.layer_input_hook(i, x2d)
            x2d = layer.forward_tokens(x2d, B, T, seed, 4 * i, wp)
        return x2d

    def forward(self, src, mask=None, src_key_padding_mask=None, is_causal=None):
        T, B, D = src.shape
        y = self.forward_tokens(src.transpose(0, 1).reshape(B * T, D).contiguous(), B, T)
        return y.view(B, T, D).transpose(0, 1)

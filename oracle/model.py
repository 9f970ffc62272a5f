"""ORACLE — test infrastructure, not product code.

Plain torch fp32 (CPU or any device) restatement of the reference model's forward pass, written
against a reference-shaped state_dict.  Autograd of this restatement is the gradient oracle.
Follows:
  ResBlock                              architecture.py:14-40
  Model.forward                         architecture.py:61-84
  TransformerEncoderLayer (post-norm)   transformer.py:43-60
  MultiHeadAttention                    transformer.py:87-112
  LearnedRelativePositionalEmbedding    transformer.py:162-297, via the closed form
      pos[b,h,q,k] = q[b,h,q,:] . E[h, k-q+99, :]   if |k-q| <= 99 (always true for T <= 100)
                   = -1e8                            otherwise
  which tests/golden/make_golden_model.py checks against the executed reference
  (pad / skew implementation) to fp32 rounding.
Pinned by tests/golden/model_golden.npz (outputs, loss and gradient fingerprints produced by
the executed reference on formula-defined weights).
"""
import math
import random

import torch
import torch.nn.functional as F

BN_EPS, BN_MOMENTUM, LN_EPS = 1e-5, 0.1, 1e-5


def _bn(x, sd, prefix, training, update_running):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        # batch statistics over (B, L); functional batch_norm updates rm/rv in place
        return F.batch_norm(x, rm if update_running else None, rv if update_running else None, w, b,
                            True, BN_MOMENTUM, BN_EPS)
    return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)


def res_block(x, sd, prefix, stride, training, update_running=True):
    """architecture.py:29-40; x: (B, C, L)."""
    h = F.conv1d(x, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], stride=stride,
                 padding=1)
    h = F.relu(_bn(h, sd, prefix + ".bn1", training, update_running))
    h = F.conv1d(h, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], padding=1)
    h = _bn(h, sd, prefix + ".bn2", training, update_running)
    r = F.conv1d(x, sd[prefix + ".residual_path.weight"], sd[prefix + ".residual_path.bias"],
                 stride=stride)
    r = _bn(r, sd, prefix + ".res_norm", training, update_running)
    return F.relu(h + r)


def attention(x, sd, prefix, n_head, dropout_p, training, max_rel=100):
    """transformer.py:87-112 with the closed-form relative-position logits. x: (T, B, D)."""
    wq, wk, wv, wo = (sd[prefix + n] for n in (".w_q", ".w_k", ".w_v", ".w_o"))
    E = sd[prefix + ".relative_positional.embeddings"].detach()[..., 0]   # (H, 2R-1, dh); no grad (F3)
    dh = wq.shape[2]
    q = torch.einsum('tbf,hfa->bhta', x, wq)
    k = torch.einsum('tbf,hfa->bhta', x, wk)
    v = torch.einsum('tbf,hfa->bhta', x, wv)
    logits = torch.einsum('bhqa,bhka->bhqk', q, k) / (dh ** 0.5)
    T = x.shape[0]
    Rl = torch.einsum('bhqa,hra->bhqr', q, E)                  # (B, H, T, 2R-1)
    ar = torch.arange(T, device=x.device)
    rel = ar[None, :] - ar[:, None] + (max_rel - 1)            # k - q + 99, shape (T, T)
    inband = (rel >= 0) & (rel <= 2 * max_rel - 2)
    idx = rel.clamp(0, 2 * max_rel - 2)
    pos = torch.gather(Rl, 3, idx[None, None].expand(Rl.shape[0], Rl.shape[1], T, T))
    pos = torch.where(inband[None, None], pos, torch.full_like(pos, -1e8))
    probs = F.softmax(logits + pos, dim=-1)
    probs = F.dropout(probs, dropout_p, training)
    o = torch.einsum('bhqk,bhka->bhqa', probs, v)
    return torch.einsum('bhta,haf->tbf', o, wo)


def encoder_layer(x, sd, prefix, n_head, dropout_p, training):
    """transformer.py:43-60 (post-norm). x: (T, B, D)."""
    a = attention(x, sd, prefix + ".self_attn", n_head, dropout_p, training)
    x = x + F.dropout(a, dropout_p, training)
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"], LN_EPS)
    h = F.relu(F.linear(x, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"]))
    h = F.linear(F.dropout(h, dropout_p, training), sd[prefix + ".linear2.weight"],
                 sd[prefix + ".linear2.bias"])
    x = x + F.dropout(h, dropout_p, training)
    return F.layer_norm(x, (D,), sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"], LN_EPS)


def num_layers_of(sd):
    return 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))


def model_forward(sd, x_raw, training=False, dropout_p=0.0, n_head=8, update_running=True,
                  shift=None):
    """architecture.py:61-84.  sd: reference-shaped state_dict (tensors may require grad).
    x_raw: (B, L, 8).  In training mode the reference shifts x_raw left by
    r = random.randrange(8) in place (:64-68); pass `shift` to force r, else Python's RNG is used.
    Returns (w_out(x), w_aux(x)) or w_out(x)."""
    if training:
        r = random.randrange(8) if shift is None else shift
        if r > 0:
            x_raw[:, :-r, :] = x_raw[:, r:, :].clone()
            x_raw[:, -r:, :] = 0
    x = x_raw.transpose(1, 2)
    for i in range(3):
        x = res_block(x, sd, f"conv_blocks.{i}", 2, training, update_running)
    x = x.transpose(1, 2)
    x = F.linear(x, sd["w_raw_in.weight"], sd["w_raw_in.bias"])
    x = x.transpose(0, 1)
    for i in range(num_layers_of(sd)):
        x = encoder_layer(x, sd, f"transformer.layers.{i}", n_head, dropout_p, training)
    x = x.transpose(0, 1)
    out = F.linear(x, sd["w_out.weight"], sd["w_out.bias"])
    if "w_aux.weight" in sd:
        return out, F.linear(x, sd["w_aux.weight"], sd["w_aux.bias"])
    return out


# ---------------------------------------------------------------------------------------
# formula-defined weights: reproducible anywhere without RNG or the reference constructor
# ---------------------------------------------------------------------------------------
def formula_state_dict(model_size, num_layers, num_outs=80, num_aux=48, n_head=8, ff=3072,
                       max_rel=100, dtype=torch.float32):
    """Deterministic reference-shaped state_dict: w = amp * sin(a*i + b) per tensor."""
    D, dh = model_size, model_size // n_head
    shapes = {}
    cin = 8
    for i in range(3):
        p = f"conv_blocks.{i}"
        shapes[p + ".conv1.weight"] = (D, cin, 3); shapes[p + ".conv1.bias"] = (D,)
        shapes[p + ".conv2.weight"] = (D, D, 3); shapes[p + ".conv2.bias"] = (D,)
        shapes[p + ".residual_path.weight"] = (D, cin, 1); shapes[p + ".residual_path.bias"] = (D,)
        for bn in ("bn1", "bn2", "res_norm"):
            for n in ("weight", "bias", "running_mean", "running_var"):
                shapes[f"{p}.{bn}.{n}"] = (D,)
            shapes[f"{p}.{bn}.num_batches_tracked"] = ()
        cin = D
    shapes["w_raw_in.weight"] = (D, D); shapes["w_raw_in.bias"] = (D,)
    for l in range(num_layers):
        p = f"transformer.layers.{l}"
        for n in ("w_q", "w_k", "w_v"):
            shapes[f"{p}.self_attn.{n}"] = (n_head, D, dh)
        shapes[f"{p}.self_attn.w_o"] = (n_head, dh, D)
        shapes[f"{p}.self_attn.relative_positional.embeddings"] = (n_head, 2 * max_rel - 1, dh, 1)
        shapes[f"{p}.linear1.weight"] = (ff, D); shapes[f"{p}.linear1.bias"] = (ff,)
        shapes[f"{p}.linear2.weight"] = (D, ff); shapes[f"{p}.linear2.bias"] = (D,)
        for n in ("norm1", "norm2"):
            shapes[f"{p}.{n}.weight"] = (D,); shapes[f"{p}.{n}.bias"] = (D,)
    shapes["w_out.weight"] = (num_outs, D); shapes["w_out.bias"] = (num_outs,)
    if num_aux is not None:
        shapes["w_aux.weight"] = (num_aux, D); shapes["w_aux.bias"] = (num_aux,)
    sd = {}
    for j, (name, shp) in enumerate(shapes.items()):
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.int64)
            continue
        n = 1
        for s in shp:
            n *= s
        i = torch.arange(n, dtype=torch.float64)
        wave = torch.sin(i * (0.37 + 0.011 * (j % 17)) + 0.5 * j)
        if name.endswith("running_var") or (("norm" in name or ".bn" in name) and name.endswith("weight")):
            t = 1.0 + 0.2 * wave
        elif name.endswith("running_mean") or name.endswith("bias"):
            t = 0.1 * wave
        elif name.endswith("embeddings"):
            t = wave * (dh ** -0.5)
        else:
            fan_in = n // shp[0] if len(shp) > 1 else n
            if ".self_attn.w_" in name:
                fan_in = shp[1]
            t = wave * (1.5 / math.sqrt(max(fan_in, 1)))
        sd[name] = t.to(dtype).reshape(shp).clone()
    return sd

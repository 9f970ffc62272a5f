"""Subprocess body of tests/test_reference_integration_gpu.py (own process: it rearranges
sys.path, defines the reference's absl flags and loads two modules named `architecture`).

What runs here is the reference's transduction_model.py, UNMODIFIED (from /root/reference or its
verbatim copy baseline/_ref), with dropin/ ahead of it on sys.path:
  1. `transduction_model.dtw_loss` (:98-157) and the training-step body (:196-212: zero_grad,
     combine_fixed_length, model(...), dtw_loss, backward, AdamW.step) for 2 iterations on the
     GPU with OUR Model / align_from_distances / data_utils, next to the same lines executed on
     the CPU with the REFERENCE's Model (architecture.py + transformer.py) and numba align.py
     from the same state_dict and batch: losses, every gradient, parameters after 2 steps;
  2. `transduction_model.test` (:33-55, eval loop with phoneme accuracy + confusion matrix);
  3. `transduction_model.train_model` (:159-227) for one epoch on the synthetic corpus mirror;
  4. `transduction_model.save_output` (:57-71) with the drop-in `Vocoder` (HiFi-GAN on libssb) against
     the reference's own generator on the CPU.
Prints one JSON object.
"""
import json
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from baseline import refenv  # noqa: E402


def main():
    D, NL = int(sys.argv[1]), int(sys.argv[2])
    out = {}
    ref = refenv.reference_dir()
    ra, rt = refenv.load_reference_model_modules()          # reference Model, aliased
    ref_align = refenv._load_as(ref, "align", "ref_align")  # reference numba DTW
    tm = refenv.import_transduction_model_with_dropin(synthetic_corpus=True)
    import architecture, align, data_utils, read_emg, transformer   # noqa: E401  (the drop-ins)
    for m in (architecture, align, data_utils, read_emg, transformer):
        assert "dropin" in m.__file__, m.__file__
    assert tm.Model is architecture.Model and tm.align_from_distances is align.align_from_distances
    assert "silent_speech_b200" in tm.Model.__module__
    out["transduction_model_file"] = tm.__file__

    from absl import flags
    FLAGS = flags.FLAGS
    tmp = tempfile.mkdtemp(prefix="ssb_ref_")
    FLAGS(["t", f"--model_size={D}", f"--num_layers={NL}", "--dropout=0.0", "--epochs=1",
           f"--output_directory={tmp}"])

    from silent_speech_b200.read_emg import EMGDataset
    torch.manual_seed(0)
    ref_model = refenv.fix_transformer_shim(ra.Model(112, 80, 48)).train()
    sd = {k: v.clone() for k, v in ref_model.state_dict().items()}
    ours = tm.Model(112, 80, 48)
    ours.load_state_dict(sd, strict=True)
    ours = ours.to("cuda").train()

    ds = EMGDataset(num_examples=10, frames=140, frames_jitter=30, seed=5)
    batch = EMGDataset.collate_raw([ds[i] for i in range(len(ds))])
    out["n_silent"] = int(sum(batch["silent"]))
    assert 2 <= out["n_silent"] <= 8

    optim_ref = torch.optim.AdamW(ref_model.parameters(), weight_decay=FLAGS.l2)
    optim_our = torch.optim.AdamW(ours.parameters(), weight_decay=FLAGS.l2)
    seq_len = 200

    def step_body(model, optim, device, it):
        """transduction_model.py:197-210, same calls in the same order."""
        random.seed(100 + it)                  # Model.forward draws the shift with `random`
        optim.zero_grad()
        X = tm.combine_fixed_length([t.to(device, non_blocking=True) for t in batch['emg']], seq_len)
        X_raw = tm.combine_fixed_length([t.to(device, non_blocking=True) for t in batch['raw_emg']],
                                        seq_len * 8)
        sess = tm.combine_fixed_length([t.to(device, non_blocking=True) for t in batch['session_ids']],
                                       seq_len)
        pred, phoneme_pred = model(X, X_raw, sess)
        loss, _ = tm.dtw_loss(pred, phoneme_pred, batch)
        lv = loss.item()
        loss.backward()
        grads = {k: p.grad.detach().cpu().double().clone() for k, p in model.named_parameters()
                 if p.grad is not None}
        optim.step()
        return lv, grads, pred.detach().cpu().double()

    losses_ref, losses_our, grad_err, pred_err = [], [], [], []
    worst = ("", 0.0)
    for it in range(2):
        saved = tm.align_from_distances
        tm.align_from_distances = ref_align.align_from_distances   # pure-reference CPU pass
        # architecture.py:64-68 shifts x_raw IN PLACE between overlapping slices; ATen's CPU copy
        # splits that across intra-op threads and the chunk boundaries then read already-shifted
        # samples (16 of 89 600 elements here, run-to-run different).  One thread = the intended
        # shift, which is what the drop-in implements.
        nthreads = torch.get_num_threads()
        torch.set_num_threads(1)
        try:
            lr_, gr, pr = step_body(ref_model, optim_ref, "cpu", it)
        finally:
            tm.align_from_distances = saved
            torch.set_num_threads(nthreads)
        lo, go, po = step_body(ours, optim_our, "cuda", it)
        losses_ref.append(lr_)
        losses_our.append(lo)
        pred_err.append(((po - pr).norm() / pr.norm()).item())
        assert set(go) == set(gr), set(go) ^ set(gr)
        w = 0.0
        for k in gr:
            if k.startswith("conv_blocks") and k.endswith(("conv1.bias", "conv2.bias",
                                                           "residual_path.bias")):
                continue          # analytically zero in front of a training-mode BatchNorm
            e = ((go[k] - gr[k]).norm() / (gr[k].norm() + 1e-30)).item()
            if e > w:
                w = e
            if e > worst[1]:
                worst = (k, e)
        grad_err.append(w)
    out.update(losses_ref=losses_ref, losses_ours=losses_our, pred_rel_l2=pred_err,
               worst_grad_rel_l2=grad_err, worst_grad_param=worst[0])
    perr = 0.0
    osd = ours.state_dict()
    for k, v in ref_model.state_dict().items():
        if k.startswith("conv_blocks") and k.endswith(("conv1.bias", "conv2.bias",
                                                       "residual_path.bias")):
            continue              # zero-gradient parameters: Adam turns rounding noise into +-lr steps
        if v.is_floating_point():
            a, b = osd[k].detach().cpu().double(), v.double()
            perr = max(perr, ((a - b).norm() / (b.norm() + 1e-30)).item())
    out["param_rel_l2_after_2_steps"] = perr
    assert any(k.endswith("relative_positional.embeddings") for k in osd)
    for k, p in ours.named_parameters():
        if k.endswith("relative_positional.embeddings"):
            assert p.grad is None

    # 2. the reference's validation loop on our model (phoneme_eval=True path, :129-137)
    dev = EMGDataset(num_examples=6, frames=120, frames_jitter=20, seed=9, dev=True)
    val, acc, conf = tm.test(ours, dev, "cuda")
    out["test_loss"], out["test_phoneme_acc"] = float(val), float(acc)
    out["confusion_total"] = float(conf.sum())
    assert ours.training                                           # test() restores train mode

    # 3. the reference's train_model for one epoch (SizeAwareSampler budget 256000 raw samples)
    train = EMGDataset(num_examples=180, frames=400, frames_jitter=40, seed=3)
    random.seed(1)
    torch.manual_seed(1)
    model = tm.train_model(train, dev, "cuda", save_sound_outputs=False)
    out["train_model_type"] = type(model).__module__ + "." + type(model).__name__
    out["model_pt_saved"] = os.path.exists(os.path.join(tmp, "model.pt"))
    state = torch.load(os.path.join(tmp, "model.pt"), map_location="cpu")
    out["saved_keys_match_reference"] = sorted(state.keys()) == sorted(sd.keys())
    out["finite_after_epoch"] = bool(all(torch.isfinite(v).all() for v in state.values()
                                         if v.is_floating_point()))
    # two SizeAwareSampler batches (~80 utterances each) were trained on: BatchNorm counted them
    out["train_batches"] = int(state["conv_blocks.0.bn1.num_batches_tracked"])
    out["trained"] = out["train_batches"] >= 1
    # 4. the reference's save_output (:57-71) with the drop-in Vocoder (vocoder.py:16-36 on libssb):
    #    model.eval() -> pred -> normaliser inverse -> vocoder -> sf.write, against the reference's own
    #    HiFi-GAN Generator (CPU) fed the same predicted mel from the same checkpoint file
    import vocoder as vocoder_mod
    assert "dropin" in vocoder_mod.__file__ and tm.Vocoder is vocoder_mod.Vocoder
    from oracle import vocoder as ov
    vcfg = dict(ov.V1, upsample_initial_channel=128)
    with open(os.path.join(tmp, "config.json"), "w") as f:
        json.dump(vcfg, f)
    ckpt = os.path.join(tmp, "g_00000001")
    torch.save({"generator": ov.weight_normed_state_dict(ov.formula_state_dict(vcfg))}, ckpt)
    FLAGS.hifigan_checkpoint = ckpt
    voc = tm.Vocoder()
    seen = {}

    def spy(y):
        seen["mel"] = y.detach().cpu().clone()
        return voc(y)
    written = {}
    tm.sf.write = lambda fn, audio, sr: written.update(fn=fn, audio=np.asarray(audio), sr=sr)
    tm.save_output(model, dev[0], os.path.join(tmp, "example_output_0.wav"), "cuda", dev.mfcc_norm, spy)
    models, env = refenv.import_reference("models", "env")
    gen = models.Generator(env.AttrDict(vcfg))
    gen.load_state_dict(torch.load(ckpt)["generator"])
    gen.eval()
    gen.remove_weight_norm()
    with torch.no_grad():
        ref_audio = gen(seen["mel"].T[np.newaxis, :, :]).squeeze().numpy()      # vocoder.py:32-36
    out["save_output_samples"] = int(written["audio"].shape[0])
    out["save_output_sr"] = int(written["sr"])
    out["save_output_frames"] = int(seen["mel"].shape[0])
    out["vocoder_rel_l2"] = float(np.linalg.norm(written["audio"].astype(np.float64) - ref_audio)
                                  / np.linalg.norm(ref_audio))
    from silent_speech_b200 import _lib
    out["libssb_launches"] = int(_lib.launch_count)
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()

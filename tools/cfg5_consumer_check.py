#!/usr/bin/env python
"""cfg5, CPU half: does the reference's consumer see a difference?

Runs HERE (build container, /root/reference present): loads gpurun_out/cfg5_feats.npz written by
tools/cfg5_features.py on the B200, recomputes the features of the same waveforms with the UNMODIFIED reference
(lid/audio_processor.py wav2mel(use_kaildi=True)), and pushes both through the reference's
ConformerMutiLangModel (random init, torch.manual_seed(0), eval, lang=None; ref: lid/ConformerLangModel.py:77-83).
Reports the feature difference and what it does to the language-id scores and the per-language CTC logits."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Metric:            # torchmetrics objects are only constructed by the model, never called here
    def __init__(self, *a, **k):
        pass


def main():
    stub("augment")
    stub("torchmetrics", WER=_Metric, CharErrorRate=_Metric, Accuracy=_Metric, WordErrorRate=_Metric)
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        stub("torch.utils.tensorboard", SummaryWriter=_Metric)
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, "/root/reference/lid")
    import lid.audio_processor as ap
    from lid.ConformerLangModel import ConformerMutiLangModel

    z = np.load(os.path.join(ROOT, "gpurun_out", "cfg5_feats.npz"))
    wavs, ours = torch.from_numpy(z["wavs"]), torch.from_numpy(z["feats"])
    ref = torch.stack([ap.wav2mel(w.unsqueeze(0), use_kaildi=True).squeeze(0).transpose(0, 1) for w in wavs])
    assert ref.shape == ours.shape, (ref.shape, ours.shape)
    d = (ours - ref).abs()
    print("features %s: max|gpu-ref| %.3e, mean %.3e, norm-rel %.3e" %
          (tuple(ours.shape), d.max(), d.mean(), d.max() / ref.abs().max()))

    torch.manual_seed(0)
    model = ConformerMutiLangModel(lang2vocab={"Persian": 40, "Swahili": 30, "Vietnamese": 90},
                                   lang2index={"Persian": 0, "Swahili": 1, "Vietnamese": 2}, conformer_linear=True,
                                   sub_sampling=2).eval()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        out_r, lid_r = model(ref, 16000, None)
        out_o, lid_o = model(ours, 16000, None)

    def flat(x):
        if isinstance(x, dict):
            return [(k, v) for k, v in x.items()]
        if isinstance(x, (list, tuple)):
            return [(str(i), v) for i, v in enumerate(x)]
        return [("", x)]

    worst = 0.0
    for (k, a), (_, b) in zip(flat(out_r), flat(out_o)):
        if torch.is_tensor(a):
            e = (a - b).abs().max().item()
            worst = max(worst, e)
            print("ctc logits[%s] %s: max|diff| %.3e (|ref|max %.3f), argmax agreement %.4f" %
                  (k, tuple(a.shape), e, a.abs().max(), (a.argmax(-1) == b.argmax(-1)).float().mean()))
    lr = lid_r[0] if isinstance(lid_r, (list, tuple)) else lid_r
    lo = lid_o[0] if isinstance(lid_o, (list, tuple)) else lid_o
    if torch.is_tensor(lr):
        e = (lr - lo).abs().max().item()
        worst = max(worst, e)
        print("lid scores %s: max|diff| %.3e (|ref|max %.3f), argmax agreement %.4f" %
              (tuple(lr.shape), e, lr.abs().max(), (lr.argmax(-1) == lo.argmax(-1)).float().mean()))
    print("worst consumer-output difference: %.3e" % worst)


if __name__ == "__main__":
    main()

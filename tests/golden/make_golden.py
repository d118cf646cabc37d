#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference in the build container.

Run once, here (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference module ``lid/audio_processor.py`` is imported from /root/reference after stubbing the
``augment`` (WavAugment) package it imports at module level but only uses for reverb
(``lid/audio_processor.py:4,155-163``).  Its arithmetic is torchaudio's (2.11.0 here; the reference
pins 0.12.1 -- equivalence not verifiable offline).  MFCC does not exist in the reference; its vectors
come from ``torchaudio.compliance.kaldi.mfcc`` called with the reference's framing arguments.

Inputs are stored next to the outputs so the fixtures do not depend on RNG reproducibility.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.modules["augment"] = types.ModuleType("augment")
sys.path.insert(0, "/root/reference")

import lid.audio_processor as ap  # noqa: E402  (the reference itself)
import torchaudio  # noqa: E402
import torchaudio.compliance.kaldi as K  # noqa: E402

from oracle import frontend_oracle as O  # noqa: E402  (only for the synthetic input generators)


def main():
    torch.set_num_threads(1)
    meta = dict(torch=torch.__version__, torchaudio=torchaudio.__version__)

    # ---- kaldi fbank (A4): white-noise + speech-like, plus the frame-count edge cases -------------
    fb = {}
    cases = [("noise_3s", O.synth_noise(48000, 0)), ("noise_1s", O.synth_noise(16000, 1)),
             ("speech_2s", O.synth_speechlike(32000, 2)),
             ("n400", O.synth_noise(400, 3)), ("n559", O.synth_noise(559, 4)),
             ("n560", O.synth_noise(560, 5)), ("n1234", O.synth_noise(1234, 6)),
             ("silence", ap.normalize_wav(torch.zeros(1, 4000))),
             ("dc_ramp", ap.normalize_wav(torch.linspace(-1, 1, 3000).unsqueeze(0) + 0.5))]
    for name, x in cases:
        fb["in_" + name] = x.numpy()
        fb["out_" + name] = ap.wav2mel(x, use_kaildi=True).numpy()      # (1, 80, T)
    np.savez_compressed(os.path.join(HERE, "fbank_kaldi.npz"), **fb)

    # ---- default branch (A9): MelSpectrogram + AmplitudeToDB(top_db=80), pad 0 and 16 (the yaml value) ------------
    ms = {}
    for name, x, pad in (("noise_1s", O.synth_noise(16000, 41), 0), ("speech_1s_pad16", O.synth_speechlike(16000, 42), 16),
                         ("n700", O.synth_noise(700, 43), 0), ("n4000_pad16", O.synth_noise(4000, 44), 16),
                         ("quiet_tail", torch.cat([O.synth_noise(8000, 45), 1e-4 * O.synth_noise(8000, 46)], 1), 0)):
        ms["in_" + name] = x.numpy()
        ms["pad_" + name] = np.array([pad])
        ms["out_" + name] = ap.wav2mel(x, use_kaildi=False, pad=pad).numpy()            # (1, 80, T)
    np.savez_compressed(os.path.join(HERE, "melspec_db.npz"), **ms)

    # ---- MFCC (A5): torchaudio kaldi.mfcc with the reference's framing args ------------------------
    mf = {}
    for name, x in (("noise_1s", O.synth_noise(16000, 11)), ("speech_1s", O.synth_speechlike(16000, 12)),
                    ("n560", O.synth_noise(560, 13))):
        mf["in_" + name] = x.numpy()
        mf["out_" + name] = K.mfcc(x, num_ceps=40, num_mel_bins=80, cepstral_lifter=22.0, dither=0.0,
                                   frame_length=25, frame_shift=10, preemphasis_coefficient=1.0,
                                   sample_frequency=16000).numpy()            # (T, 40)
    np.savez_compressed(os.path.join(HERE, "mfcc_kaldi.npz"), **mf)

    # ---- SpecAugment (A6): reference outputs + the RNG-derived integer bounds ----------------------
    sa = {}
    for name, n, seed, kw in (("t798_default", 128000, 1234, dict(t_mask=0.05, f_mask=27, mask_times=2)),
                              ("t298_yaml", 48000, 77, dict(t_mask=0.05, f_mask=12, mask_times=1)),
                              ("t18_no_tmask", 3300, 5, dict(t_mask=0.05, f_mask=27, mask_times=2)),
                              ("t98_bigf", 16000, 9, dict(t_mask=0.05, f_mask=100, mask_times=1)),
                              ("t98_off", 16000, 9, dict(t_mask=0.05, f_mask=27, mask_times=0))):
        x = O.synth_noise(n, seed)
        spec = ap.wav2mel(x, use_kaildi=True)
        torch.manual_seed(seed)
        try:
            out = ap.spectrogram_augment(spec, **kw).numpy()
        except ValueError:
            out = np.zeros(0, dtype=np.float32)            # f_mask > n_mels can make torchaudio raise
        sa["spec_" + name] = spec.numpy()
        sa["out_" + name] = out
        sa["kw_" + name] = np.array([kw["t_mask"], kw["f_mask"], kw["mask_times"], seed], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "specaug.npz"), **sa)

    # ---- waveform-level stages (A1, A2) -------------------------------------------------------------
    wv = {}
    g = torch.Generator().manual_seed(21)
    raw = (torch.randn(1, 20000, generator=g) * 0.1 + 0.03)
    wv["raw"] = raw.numpy()
    wv["normalized"] = ap.normalize_wav(raw).numpy()
    torch.manual_seed(22)
    noise = torch.rand_like(raw)                              # the draw wav_augment will make
    torch.manual_seed(22)
    aug, _ = ap.wav_augment(raw.clone(), 16000)               # mutates its input: pass a clone
    wv["dither_noise"] = noise.numpy()
    wv["augmented"] = aug.numpy()
    np.savez_compressed(os.path.join(HERE, "waveform_stages.npz"), **wv)

    # ---- feature contract (A8): three ragged utterances through wav2mel -> collate-style padding -----
    ct = {}
    specs = []
    for i, n in enumerate((16000, 9000, 12345)):
        x = O.synth_noise(n, 30 + i)
        ct["in_%d" % i] = x.numpy()
        specs.append(ap.wav2mel(x, use_kaildi=True))
    # ref: lid/raw_datasets.py:345-365 (collate_fn), restated inline because raw_datasets imports ccml
    rows = [s.squeeze(0).transpose(0, 1) for s in specs]
    wavs = torch.nn.utils.rnn.pad_sequence(rows, batch_first=True)
    ct["wavs"] = wavs.numpy()
    ct["wav_percents"] = np.array([r.shape[0] / wavs.shape[1] for r in rows], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "collate.npz"), **ct)

    # ---- resampling DataProcessor (row f4): the reference's own module on ragged 44.1 / 22.05 kHz batches ----------
    rs = {}
    sys.modules.setdefault("torchmetrics", types.ModuleType("torchmetrics"))
    for n in ("WER", "CharErrorRate", "Accuracy", "WordErrorRate"):
        setattr(sys.modules["torchmetrics"], n, type(n, (), {"__init__": lambda self, *a, **k: None}))
    sys.path.insert(0, "/root/reference/lid")
    from lid.ConformerLangModel import DataProcessor            # noqa: E402  (ref: lid/ConformerLangModel.py:131-178)
    dp = DataProcessor(16000)
    g = torch.Generator().manual_seed(61)
    for rate, lens in ((44100, (22050, 9001, 13333)), (22050, (11025, 4000, 7777, 300))):
        xs = [torch.randn(n, generator=g) * 0.3 for n in lens]
        ys = dp(xs, rate)
        for i, (x, y) in enumerate(zip(xs, ys)):
            rs["in_%d_%d" % (rate, i)] = x.numpy()
            rs["out_%d_%d" % (rate, i)] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "resample.npz"), **rs)

    # ---- the wav2vec-exp FBank variant (row f4): the reference's own class, cut out of its module by name (importing
    #      wav2vec-exp/s3prl_model.py whole pulls in s3prl and fairseq, which are not installed) ---------------------------
    import ast
    src = open("/root/reference/wav2vec-exp/s3prl_model.py").read()
    node = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "FBank"][0]
    ns = {"torch": torch, "nn": torch.nn}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "wav2vec-exp/s3prl_model.py", "exec"), ns)
    FBank = ns["FBank"]
    fbk = {}
    for name, n_fft, n in (("a", 640, 640), ("b", 640, 16000), ("c", 640, 48017), ("d", 320, 8000), ("e", 640, 959)):
        xw = O.synth_noise(n, 900 + n) if name != "c" else O.synth_speechlike(n, 901)
        fbk["in_" + name] = xw.numpy()
        fbk["nfft_" + name] = np.int64(n_fft)
        fbk["out_" + name] = FBank(80, n_fft)(xw).numpy()
    np.savez_compressed(os.path.join(HERE, "s3prl_fbank.npz"), **fbk)

    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as f:
        f.write("generated by tests/golden/make_golden.py from /root/reference (kouyt5/speech-lid)\n")
        for k, v in meta.items():
            f.write("%s %s\n" % (k, v))
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()

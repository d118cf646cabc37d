#!/usr/bin/env python
"""SURVEY.md 8(c) acceptance metrics (ii) and (iv), exactly as written, for every BASELINE config -- pass / fail COUNTS.

  (ii)  per utterance:  max|gpu - oracle32| / max|oracle32| <= 1e-4   over ALL output dims
  (iv)  per utterance and per output dim (max over frames):  |gpu - truth64| <= 1.5 x |oracle32 - truth64|
(the test-suite asserts a relaxed form of both, see tests/test_gpu_parity.py; this table is the unrelaxed statement.)
Also printed: the same two metrics for an INDEPENDENT CPU fp32 implementation of the reference's arithmetic (identical
fp32 frames and tables, numpy's pocketfft instead of torch's FFT) against the oracle, i.e. how two CPU libraries fare
under the same rule.  Run on the GPU box: python tools/parity_matrix.py [--quick] [--precise]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import speech_lid_b200 as lid
from oracle import frontend_oracle as O

quick = "--quick" in sys.argv
precise = "--precise" in sys.argv      # FrontEnd(precise=True): the float64 kernel (lidfe_fbank_precise.cuh)
torch.set_num_threads(max(1, (os.cpu_count() or 8) // 2))


def cpu_other_fft(x, n_mels=80):
    frames = O.kaldi_windowed_frames(x[0].to(torch.float32), 400, 160, 512, 1.0)
    X = torch.from_numpy(np.fft.rfft(frames.numpy()).astype(np.complex64))
    mel = torch.mm(X.abs().pow(2.0), O.kaldi_mel_banks(n_mels, 512, 16000.0).T)
    return torch.max(mel, torch.tensor(O.EPS32)).log()


def truth_mfcc(x):
    f = O.truth64_fbank(x)
    dct = O.kaldi_dct_matrix(40, 80).double()
    return (f @ dct) * O.kaldi_lifter(40, 22.0).double()


GROUPS = ((0, 3), (3, 10), (10, 40), (40, 80))
group_fail = {}


def score(got, ref, tru, tag=None):
    ii = float((got - ref).abs().max() / ref.abs().max())
    eg = (got.double() - tru).abs().max(0).values
    er = (ref.double() - tru).abs().max(0).values
    ratio = eg / er.clamp_min(1e-30)
    if tag is not None:
        for a, b in GROUPS:
            if a < ratio.numel():
                k = (tag, a, b)
                group_fail[k] = group_fail.get(k, 0) + int((ratio[a:b] > 1.5).sum())
    return ii, float(ratio.max()), int((ratio > 1.5).sum())


def run(name, wavs, fe, oracle, truth, other=None):
    feats, _ = fe.featurize(wavs)
    feats = feats.cpu()
    acc = dict(utts=len(wavs), ii_fail=0, ii_worst=0.0, iv_fail_utts=0, iv_fail_dims=0, iv_worst=0.0, dims=0)
    oth = dict(ii_fail=0, ii_worst=0.0, iv_fail_utts=0, iv_fail_dims=0, iv_worst=0.0)
    for i, w in enumerate(wavs):
        ref = oracle(w)
        tru = truth(w)
        got = feats[i, :ref.shape[0]]
        ii, worst, nd = score(got, ref, tru, name)
        acc["ii_fail"] += ii > 1e-4; acc["ii_worst"] = max(acc["ii_worst"], ii)
        acc["iv_fail_utts"] += nd > 0; acc["iv_fail_dims"] += nd; acc["iv_worst"] = max(acc["iv_worst"], worst)
        acc["dims"] += ref.shape[1]
        if other is not None:
            ii, worst, nd = score(other(w), ref, tru)
            oth["ii_fail"] += ii > 1e-4; oth["ii_worst"] = max(oth["ii_worst"], ii)
            oth["iv_fail_utts"] += nd > 0; oth["iv_fail_dims"] += nd; oth["iv_worst"] = max(oth["iv_worst"], worst)
    line = ("%-34s utts %4d | (ii) fail %4d worst %.3g | (iv) fail utts %4d, dims %5d of %6d, worst ratio %.2f" % (
        name, acc["utts"], acc["ii_fail"], acc["ii_worst"], acc["iv_fail_utts"], acc["iv_fail_dims"], acc["dims"], acc["iv_worst"]))
    print(line, flush=True)
    print("%-34s           | (iv) failing dims by output-dim group: %s" % ("", ", ".join("%d-%d: %d" % (a, b - 1, group_fail.get((name, a, b), 0)) for a, b in GROUPS)), flush=True)
    if other is not None:
        print("%-34s           | (ii) fail %4d worst %.3g | (iv) fail utts %4d, dims %5d of %6d, worst ratio %.2f" % (
            "   numpy pocketfft f32 (CPU) vs oracle", oth["ii_fail"], oth["ii_worst"], oth["iv_fail_utts"], oth["iv_fail_dims"], acc["dims"], oth["iv_worst"]), flush=True)
    return dict(gpu=acc, cpu_other=oth if other is not None else None)


def main():
    fe = lid.FrontEnd(n_mels=80, precise=precise)
    fem = lid.FrontEnd(n_mels=80, n_ceps=40, precise=precise)
    print("arithmetic: %s" % ("precise (float64 kernel)" if precise else "fast (fp32 kernels)"), flush=True)
    res = {}
    n2, n3, n4 = (32, 32, 32) if quick else (256, 512, 256)
    g = torch.Generator().manual_seed(3)
    lens4 = torch.randint(16000, 320001, (10000,), generator=g).tolist()[:n4]
    res["cfg1"] = run("cfg1 32 x 3 s noise, fbank", [O.synth_noise(48000, s) for s in range(32)], fe, O.kaldi_fbank, O.truth64_fbank, cpu_other_fft)
    res["cfg2"] = run("cfg2 %d x 8 s noise, fbank" % n2, [O.synth_noise(128000, 100 + s) for s in range(n2)], fe, O.kaldi_fbank, O.truth64_fbank, cpu_other_fft)
    res["cfg2_speech"] = run("cfg2 %d x 8 s speech-like, fbank" % min(n2, 64), [O.synth_speechlike(128000, 200 + s) for s in range(min(n2, 64))], fe, O.kaldi_fbank, O.truth64_fbank, cpu_other_fft)
    res["cfg3"] = run("cfg3 %d x 4 s noise, MFCC-40" % n3, [O.synth_noise(64000, 300 + s) for s in range(n3)], fem, O.kaldi_mfcc, truth_mfcc)
    res["cfg4"] = run("cfg4 %d x 1-20 s noise, fbank" % n4, [O.synth_noise(n, 400 + s) for s, n in enumerate(lens4)], fe, O.kaldi_fbank, O.truth64_fbank)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/parity_matrix%s.json" % ("_precise" if precise else ""), "w"), indent=1)


if __name__ == "__main__":
    main()

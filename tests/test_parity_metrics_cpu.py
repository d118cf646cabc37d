"""CPU-only facts about SURVEY.md 8(c)'s acceptance metrics that DESIGN.md section 2 leans on (no GPU involved):

  * an evaluation of the reference's formula that is EXACT up to the final rounding (float64 throughout on the reference's
    fp32 tables -- what fbank_precise_kernel computes on the device) meets metric (iv) with room to spare;
  * the same exact evaluation "fails" metric (ii) (max|x - oracle32| / max|oracle32| <= 1e-4 over all bins) on some
    white-noise utterances, because for it the metric IS the reference's own fp32 error against the truth: (ii) cannot be
    met on every utterance by computing the formula correctly, only by reproducing the oracle's FFT round-off.
"""
import torch

from oracle import frontend_oracle as O


def _exact_rounded_once(w):
    return O.truth64_fbank(w).to(torch.float32)


def test_exact_evaluation_meets_metric_iv_and_inherits_metric_ii():
    worst_ii = 0.0
    for s in range(16):
        w = O.synth_noise(128000, 100 + s)          # the inputs of the GPU suite's strict tests
        ref, tru = O.kaldi_fbank(w), O.truth64_fbank(w)
        got = _exact_rounded_once(w)
        eg = (got.double() - tru).abs().max(0).values
        er = (ref.double() - tru).abs().max(0).values
        assert bool((eg <= 1.5 * er).all())                                   # (iv), as written
        assert float(eg.max()) <= 1.0e-6                                      # half an ulp of features up to 16
        mine = float((got - ref).abs().max() / ref.abs().max())               # (ii) of the exact result ...
        theirs = float((ref.double() - tru).abs().max() / ref.abs().max())    # ... is the reference's own error
        assert abs(mine - theirs) <= 2e-7
        worst_ii = max(worst_ii, mine)
    # on this host's torch FFT the reference itself is more than 1e-4 of the range away from the truth on at least one of
    # the 16 utterances (2.9e-4 in the build container and on the GPU box); if a future torch build is more accurate
    # than that, this line -- and DESIGN.md section 2 -- should be revisited
    assert worst_ii > 1e-4, worst_ii

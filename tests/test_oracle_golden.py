"""The oracle (oracle/frontend_oracle.py) against the reference's own outputs (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference) and, when importable, against torchaudio live.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as O


def _npz(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _close_or_equal(a, b, what):
    """Same torch build -> bit-exact.  A different CPU/BLAS (the GPU box's host) may pick other vector kernels,
    so fall back to 1e-5 abs (fp32 log-mel of magnitude ~10) before failing."""
    if torch.equal(a, b):
        return
    d = (a - b).abs().max().item()
    assert d <= 2e-5 * max(1.0, b.abs().max().item()), "%s differs from the reference fixture by %g" % (what, d)


def test_fbank_matches_reference_fixtures(golden_dir):
    z = _npz(golden_dir, "fbank_kaldi.npz")
    names = [k[3:] for k in z.files if k.startswith("in_")]
    assert len(names) >= 8
    for name in names:
        x = torch.from_numpy(z["in_" + name])
        want = torch.from_numpy(z["out_" + name])
        got = O.wav2mel_kaldi(x)
        assert got.shape == want.shape, name
        assert got.shape[-1] == O.kaldi_num_frames(x.shape[-1])
        _close_or_equal(got, want, "fbank " + name)


def test_frame_counts():
    # ta: compliance/kaldi.py:63-67; values verified against the reference in SURVEY.md §8
    for n, m in ((400, 1), (559, 1), (560, 2), (16000, 98), (48000, 298), (64000, 398), (128000, 798),
                 (320000, 1998), (399, 0), (0, 0)):
        assert O.kaldi_num_frames(n) == m


def test_short_utterance_asserts():
    with pytest.raises(AssertionError):
        O.wav2mel_kaldi(torch.zeros(1, 399))


def test_silence_is_log_eps(golden_dir):
    z = _npz(golden_dir, "fbank_kaldi.npz")
    want = torch.from_numpy(z["out_silence"])
    assert torch.all(want == want.flatten()[0])
    assert abs(float(want.flatten()[0]) - (-15.942385)) < 1e-5
    got = O.wav2mel_kaldi(torch.from_numpy(z["in_silence"]))
    assert torch.equal(got, want)


def test_mfcc_matches_torchaudio_fixtures(golden_dir):
    z = _npz(golden_dir, "mfcc_kaldi.npz")
    for name in [k[3:] for k in z.files if k.startswith("in_")]:
        got = O.kaldi_mfcc(torch.from_numpy(z["in_" + name]))
        want = torch.from_numpy(z["out_" + name])
        assert got.shape == want.shape
        _close_or_equal(got, want, "mfcc " + name)


def test_specaug_matches_reference_fixtures(golden_dir):
    z = _npz(golden_dir, "specaug.npz")
    for name in [k[5:] for k in z.files if k.startswith("spec_")]:
        spec = torch.from_numpy(z["spec_" + name])
        t_mask, f_mask, mask_times, seed = z["kw_" + name]
        want = z["out_" + name]
        torch.manual_seed(int(seed))
        if want.size == 0:       # the reference raised ValueError for this draw
            with pytest.raises(ValueError):
                O.spectrogram_augment(spec, float(t_mask), int(f_mask), int(mask_times))
            continue
        got = O.spectrogram_augment(spec, float(t_mask), int(f_mask), int(mask_times))
        assert torch.equal(got, torch.from_numpy(want)), name
        # integer bounds re-applied -> identical (bit-exact mask contract)
        torch.manual_seed(int(seed))
        bounds = O.draw_mask_bounds(spec.shape[2], spec.shape[1], float(t_mask), int(f_mask), int(mask_times))
        assert torch.equal(O.apply_mask_bounds(spec, bounds), torch.from_numpy(want))
        if name == "t18_no_tmask":
            assert all(b[0] == 0 and b[1] == 0 for b in bounds)     # int(18*0.05) == 0 -> no time mask, no draw


def test_waveform_stages_match_reference(golden_dir):
    z = _npz(golden_dir, "waveform_stages.npz")
    raw = torch.from_numpy(z["raw"])
    _close_or_equal(O.normalize_wav(raw), torch.from_numpy(z["normalized"]), "normalize_wav")
    got = O.wav_dither_preemph(raw, torch.from_numpy(z["dither_noise"]))
    _close_or_equal(got, torch.from_numpy(z["augmented"]), "wav_augment")


def test_collate_contract(golden_dir):
    z = _npz(golden_dir, "collate.npz")
    specs = [O.wav2mel_kaldi(torch.from_numpy(z["in_%d" % i])) for i in range(3)]
    wavs, percents = O.collate_features(specs)
    _close_or_equal(wavs, torch.from_numpy(z["wavs"]), "collate wavs")
    assert torch.equal(percents, torch.from_numpy(z["wav_percents"]))
    # lengths recovered the way the consumer does (ref: lid/LidModule_ASR_Supervised.py:165).  fp32 T_i/T_max * T_max
    # can land one below T_i (54/98*98 -> 53.99999 -> 53): a quirk of the reference contract we keep, not fix.
    rec = (wavs.shape[1] * percents).long().tolist()
    assert all(t - 1 <= r <= t for r, t in zip(rec, [s.shape[2] for s in specs]))


def test_oracle_vs_torchaudio_live():
    K = pytest.importorskip("torchaudio.compliance.kaldi")
    for n, seed in ((16000, 1), (4321, 2), (400, 3)):
        x = O.synth_noise(n, seed)
        ref = K.fbank(x, num_mel_bins=80, dither=0.0, frame_length=25, frame_shift=10,
                      preemphasis_coefficient=1.0, sample_frequency=16000)
        assert torch.equal(O.kaldi_fbank(x), ref)
        ref = K.mfcc(x, num_ceps=40, num_mel_bins=80, dither=0.0, frame_length=25, frame_shift=10,
                     preemphasis_coefficient=1.0, sample_frequency=16000)
        assert torch.equal(O.kaldi_mfcc(x), ref)
    # the 0.97 coefficient (Kaldi's own default) follows the same restatement
    x = O.synth_speechlike(8000, 4)
    ref = K.fbank(x, num_mel_bins=80, dither=0.0, preemphasis_coefficient=0.97, sample_frequency=16000)
    assert torch.equal(O.kaldi_fbank(x, preemph=0.97), ref)
    # every window of _feature_window_function (ta: compliance/kaldi.py:86-113), with and without DC removal
    for w in ("povey", "hanning", "hamming", "rectangular", "blackman"):
        ref = K.fbank(x, num_mel_bins=80, dither=0.0, preemphasis_coefficient=1.0, window_type=w, sample_frequency=16000)
        assert torch.equal(O.kaldi_fbank(x, window_type=w), ref), w
    ref = K.fbank(x, num_mel_bins=80, dither=0.0, preemphasis_coefficient=0.97, window_type="hamming", remove_dc_offset=False)
    assert torch.equal(O.kaldi_fbank(x, preemph=0.97, window_type="hamming", remove_dc=False), ref)


def test_truth64_calibration():
    """fp32 oracle vs an fp64 evaluation with the same fp32 tables: the low mel bins are cancellation
    dominated with preemph=1.0 on white noise -- this is what bounds any fp32 implementation."""
    x = O.synth_noise(32000, 5)
    d = (O.kaldi_fbank(x).double() - O.truth64_fbank(x)).abs().max(0).values
    assert d[10:].max() < 2e-4
    assert d[:3].max() < 5e-3


def test_cmvn_definition():
    f = O.kaldi_fbank(O.synth_noise(16000, 6))
    y = O.cmvn_per_utt(f)
    assert y.mean(0).abs().max() < 1e-5 and (y.std(0) - 1).abs().max() < 1e-5
    feats = [O.kaldi_fbank(O.synth_noise(n, 7 + i)) for i, n in enumerate((8000, 12000, 5000))]
    mean, std = O.cmvn_finalize(O.cmvn_stats(feats))
    allf = torch.cat(feats).double()
    assert torch.allclose(mean, allf.mean(0), atol=1e-10) and torch.allclose(std, allf.std(0), atol=1e-9)


def test_oracle_data_processor_matches_reference_resampling(golden_dir):
    """Row f4: DataProcessor.forward of the reference (ref: lid/ConformerLangModel.py:131-178) on ragged 44.1 / 22.05 kHz
    batches; the oracle restates torchaudio's polyphase sinc resampler and the reference's crop rule."""
    z = np.load(os.path.join(golden_dir, "resample.npz"))
    for rate, n in ((44100, 3), (22050, 4)):
        xs = [torch.from_numpy(z["in_%d_%d" % (rate, i)]) for i in range(n)]
        ys = O.data_processor(xs, rate)
        for i in range(n):
            want = torch.from_numpy(z["out_%d_%d" % (rate, i)])
            assert ys[i].shape == want.shape, (rate, i)
            assert torch.allclose(ys[i], want, rtol=0, atol=1e-6), (rate, i, (ys[i] - want).abs().max())
    x = torch.randn(2, 3000)
    assert torch.equal(O.data_processor([x[0], x[1]], 16000)[0], x[0])         # other rates pass through untouched


def test_oracle_s3prl_fbank_matches_reference_module(golden_dir):
    """Row f4: the wav2vec-exp FBank variant (ref: wav2vec-exp/s3prl_model.py:174-204).  The fixtures are outputs of the
    reference's own class (tests/golden/make_golden.py cuts it out of its module by name); the restatement is bit-equal."""
    z = np.load(os.path.join(golden_dir, "s3prl_fbank.npz"))
    for k in "abcde":
        x = torch.from_numpy(z["in_" + k])
        want = torch.from_numpy(z["out_" + k])
        n_fft = int(z["nfft_" + k])
        got = O.s3prl_fbank(x, 80, n_fft)
        assert got.shape == want.shape and got.shape[-1] == O.s3prl_fbank_num_frames(x.shape[-1], n_fft)
        assert torch.equal(got, want), k

"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the MFCC arithmetic the reference reaches through third-party
packages at Envs/audioLoader.py:147-164 (`get_mfcc`) and :241-252
(`processSoundFeat`).

* `mfcc_torchaudio`  restates torchaudio.transforms.MFCC as configured at
  Envs/audioLoader.py:150-157 (torchaudio pinned ~=0.12.1 in requirements.txt:11;
  2.11.0 installed here): Spectrogram(power=2, center, reflect) ->
  MelScale(htk, norm=None) -> log(x + 1e-6) -> DCT-II ortho -> transpose.
  PINNED: checked against the installed torchaudio and against the reference's
  own `audioLoader.get_mfcc` by oracle/make_golden.py (tests/golden/mfcc_*.npz).
* `mfcc_psf` restates python_speech_features==0.6 `mfcc` (requirements.txt:13)
  as called at Envs/audioLoader.py:159-161.  The package is absent from this
  image: PARITY UNPINNED for this variant.
"""
import math

import numpy as np

EPS_LOG = 1e-6


def hamming_periodic(n):
    """torch.hamming_window(n) (periodic=True): 0.54 - 0.46 cos(2 pi k / n)."""
    k = np.arange(n, dtype=np.float64)
    return (0.54 - 0.46 * np.cos(2.0 * np.pi * k / n)).astype(np.float32)


def padded_window(n_fft, win_length):
    """torch.stft pads the window to n_fft, centred."""
    w = np.zeros(n_fft, dtype=np.float32)
    left = (n_fft - win_length) // 2
    w[left:left + win_length] = hamming_periodic(win_length)
    return w


def melscale_fbanks_htk(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk'), float32 like torch."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs, dtype=np.float32)
    m_min = np.float32(2595.0 * math.log10(1.0 + f_min / 700.0))
    m_max = np.float32(2595.0 * math.log10(1.0 + f_max / 700.0))
    m_pts = np.linspace(m_min, m_max, n_mels + 2, dtype=np.float32)
    f_pts = (700.0 * (np.power(np.float32(10.0), m_pts / np.float32(2595.0)) - 1.0)).astype(np.float32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(np.float32(0), np.minimum(down, up)).astype(np.float32)  # [n_freqs, n_mels]


def create_dct_ortho(n_mfcc, n_mels):
    """torchaudio.functional.create_dct(norm='ortho') -> [n_mels, n_mfcc]."""
    n = np.arange(n_mels, dtype=np.float32)
    k = np.arange(n_mfcc, dtype=np.float32)[:, None]
    dct = np.cos(np.float32(math.pi / n_mels) * (n + np.float32(0.5)) * k).astype(np.float32)
    dct[0] *= np.float32(1.0 / math.sqrt(2.0))
    dct *= np.float32(math.sqrt(2.0 / n_mels))
    return dct.T.copy()


def stft_params(fs, n_fft, win_len_time, win_step_time):
    """Envs/audioLoader.py:151-152: int(windowLenTime*fs), int(windowStepTime*fs)."""
    return n_fft, int(win_len_time * fs), int(win_step_time * fs)


def num_frames_torchaudio(n_samples, hop):
    return 1 + n_samples // hop


def mfcc_torchaudio(samples, fs=16000, n_fft=512, win_length=400, hop=160, n_mfcc=40, n_mels=40):
    """int16 or float wav [S] -> float32 [T, n_mfcc] (already transposed as audioLoader.py:157)."""
    x = np.asarray(samples)
    if x.dtype == np.int16:
        x = (x / 32768.0).astype(np.float32)  # audioLoader.py:154-155
    x = x.astype(np.float32)
    pad = n_fft // 2
    xp = np.pad(x, (pad, pad), mode="reflect")
    T = num_frames_torchaudio(len(x), hop)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    frames = xp[idx] * padded_window(n_fft, win_length)[None, :]
    spec = np.fft.rfft(frames.astype(np.float32), axis=1)
    power = (spec.real.astype(np.float32) ** 2 + spec.imag.astype(np.float32) ** 2).astype(np.float32)
    fb = melscale_fbanks_htk(n_fft // 2 + 1, 0.0, float(fs // 2), n_mels, fs)
    mel = power @ fb
    logmel = np.log(mel + np.float32(EPS_LOG)).astype(np.float32)
    return (logmel @ create_dct_ortho(n_mfcc, n_mels)).astype(np.float32)


def process_sound_feat(feat, sound_dim):
    """Envs/audioLoader.py:241-252: add leading dim, crop to F frames or zero-pad (float64 zeros)."""
    feat = np.expand_dims(feat, axis=0)
    nf = feat.shape[1]
    F = sound_dim[1]
    if F < nf:
        return feat[:, :F, :]
    pad_shape = list(sound_dim)
    pad_shape[1] = F - nf
    return np.concatenate((feat, np.zeros(pad_shape)), axis=1)


# ---------------------------------------------------------------------------
# python_speech_features==0.6 flavour (iTHOR path: audioLoader.py:159-161, 203-237)
# ---------------------------------------------------------------------------
def _round_half_up(x):
    return int(math.floor(x + 0.5))


def psf_filterbanks(nfilt, nfft, samplerate, lowfreq=0.0, highfreq=None):
    highfreq = highfreq or samplerate / 2
    lowmel = 2595.0 * np.log10(1 + lowfreq / 700.0)
    highmel = 2595.0 * np.log10(1 + highfreq / 700.0)
    melpoints = np.linspace(lowmel, highmel, nfilt + 2)
    bins = np.floor((nfft + 1) * (700.0 * (10 ** (melpoints / 2595.0) - 1)) / samplerate)
    fbank = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fbank[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fbank[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fbank


def num_frames_psf(n_samples, frame_len, frame_step):
    if n_samples <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * n_samples - frame_len) / frame_step))


def dct2_ortho(x):
    """scipy.fftpack.dct(type=2, norm='ortho', axis=1) written out."""
    n = x.shape[1]
    k = np.arange(n)[:, None]
    m = np.arange(n)[None, :]
    mat = np.cos(np.pi * (2 * m + 1) * k / (2.0 * n)) * math.sqrt(2.0 / n)
    mat[0] *= 1.0 / math.sqrt(2.0)
    return x @ mat.T


def mfcc_psf(signal, samplerate=16000, winlen=0.025, winstep=0.01, numcep=40, nfilt=40, nfft=512,
             preemph=0.97, ceplifter=22):
    """float64 [T, numcep]; winfunc=np.hamming (symmetric), appendEnergy=True."""
    sig = np.asarray(signal)
    sig = np.append(sig[0], sig[1:] - preemph * sig[:-1]).astype(np.float64)
    frame_len = _round_half_up(winlen * samplerate)
    frame_step = _round_half_up(winstep * samplerate)
    nfr = num_frames_psf(len(sig), frame_len, frame_step)
    padlen = (nfr - 1) * frame_step + frame_len
    padsig = np.concatenate((sig, np.zeros(padlen - len(sig))))
    idx = np.arange(frame_len)[None, :] + frame_step * np.arange(nfr)[:, None]
    frames = padsig[idx] * np.hamming(frame_len)[None, :]
    pspec = (1.0 / nfft) * np.square(np.absolute(np.fft.rfft(frames, nfft)))
    energy = np.sum(pspec, 1)
    energy = np.where(energy == 0, np.finfo(float).eps, energy)
    fb = psf_filterbanks(nfilt, nfft, samplerate)
    feat = pspec @ fb.T
    feat = np.where(feat == 0, np.finfo(float).eps, feat)
    feat = np.log(feat)
    feat = dct2_ortho(feat)[:, :numcep]
    n = np.arange(numcep)
    feat = (1 + (ceplifter / 2.0) * np.sin(np.pi * n / ceplifter)) * feat
    feat[:, 0] = np.log(energy)
    return feat

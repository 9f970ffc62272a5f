"""Mirror of the hot-path part of the reference's data_utils.py on the B200 kernels.

Same names / argument meaning as the reference:
  mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)
      data_utils.py:39-62   -> csrc/mel.cu through ssb_mel_fwd (no CPU path)
  dynamic_range_compression_torch / spectral_normalize_torch   data_utils.py:29-34 (fused in-kernel)
  combine_fixed_length / decollate_tensor                       data_utils.py:158-178 (views + cat)
  phoneme_inventory, FeatureNormalizer, TextTransform           host-side interface pieces
The hand-crafted EMG features, TextGrid reader, audio file IO etc. (data_utils.py:19-27,
85-156, 180-241) are CPU dataset preparation outside the hot path and are
not rebuilt here.
"""
import string

import numpy as np
import torch

from . import _lib

phoneme_inventory = ['aa', 'ae', 'ah', 'ao', 'aw', 'ax', 'axr', 'ay', 'b', 'ch', 'd', 'dh', 'dx',
                     'eh', 'el', 'em', 'en', 'er', 'ey', 'f', 'g', 'hh', 'hv', 'ih', 'iy', 'jh',
                     'k', 'l', 'm', 'n', 'nx', 'ng', 'ow', 'oy', 'p', 'r', 's', 'sh', 't', 'th',
                     'uh', 'uw', 'v', 'w', 'y', 'z', 'zh', 'sil']


# ---------------------------------------------------------------------------------------
# Slaney mel filterbank == librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its
# defaults (htk=False, norm='slaney'); the reference obtains it at data_utils.py:47.
# librosa is a third-party dependency absent from the reference checkout (environment.yml:17,
# unpinned); this follows its published algorithm.  Host-side constant setup, cached like the
# reference caches `mel_basis`.
# ---------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_part, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """(n_mels, n_fft//2+1) float32 triangular filters, Slaney scale and area normalisation."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


mel_basis = {}    # same cache names as the reference (data_utils.py:36-37)
hann_window = {}  # kept for interface parity; the window is generated inside the kernel


def _basis_for(sampling_rate, n_fft, num_mels, fmin, fmax, device):
    key = f"{sampling_rate}_{n_fft}_{num_mels}_{fmin}_{fmax}_{device}"
    ent = mel_basis.get(key)
    if ent is None:
        w = slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)
        nz = w != 0
        begin = np.where(nz.any(1), nz.argmax(1), 0).astype(np.int32)
        end = np.where(nz.any(1), w.shape[1] - nz[:, ::-1].argmax(1), 0).astype(np.int32)
        ent = (torch.from_numpy(w).to(device), torch.from_numpy(begin).to(device),
               torch.from_numpy(end).to(device))
        mel_basis[key] = ent
    return ent


def dynamic_range_compression_torch(x, C=1, clip_val=1e-5):
    """data_utils.py:29-30 (elementwise; the mel kernel fuses the C=1 case)."""
    return torch.log(torch.clamp(x, min=clip_val) * C)


def spectral_normalize_torch(magnitudes):
    """data_utils.py:32-34."""
    return dynamic_range_compression_torch(magnitudes)


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax,
                    center=False):
    """Log-mel spectrogram of a batch of clips: (B, S) fp32 -> (B, num_mels, frames) fp32.

    Mirrors data_utils.py:39-62, including the out-of-range prints (:40-43).  `y` must be a
    CUDA tensor: the arithmetic runs in csrc/mel.cu only.
    """
    if center:
        raise _lib.SSBError(-3, "mel_spectrogram: center=True is never used by the reference "
                                "(data_utils.py:79) and is not built")
    _lib.require_cuda(y, "y")
    if y.dim() != 2:
        raise ValueError("y must be (batch, samples)")
    flush_range_warnings()          # report finished checks of earlier calls (never blocks)
    lib = _lib.load()
    y = y.to(torch.float32)
    if y.stride(1) != 1:
        y = y.contiguous()
    B, S = y.shape
    basis, begin, end = _basis_for(sampling_rate, n_fft, num_mels, fmin, fmax, y.device)
    frames = lib.ssb_mel_num_frames(S, n_fft, hop_size)
    if frames < 0:
        raise _lib.SSBError(frames, "mel_spectrogram: bad sizes")
    out = torch.empty((B, num_mels, frames), dtype=torch.float32, device=y.device)
    # data_utils.py:40-43 prints a warning when samples leave [-1, 1].  The kernel tracks min / max
    # of what it reads (two atomics per warp); the two words come back by an asynchronous copy and
    # are reported by flush_range_warnings() - at the next call, or right after the caller's own
    # synchronisation (load_audio's .cpu()) - instead of a reduction pass + host read per call.
    cell = torch.zeros(2, dtype=torch.int32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(lib.ssb_mel_fwd(y.data_ptr(), B, S, y.stride(0), n_fft, hop_size, win_size,
                                   basis.data_ptr(), begin.data_ptr(), end.data_ptr(), num_mels,
                                   1e-5, out.data_ptr(), cell.data_ptr(), _lib.current_stream()))
        if B and frames and not torch.cuda.is_current_stream_capturing():
            host = torch.empty(2, dtype=torch.int32).pin_memory()
            host.copy_(cell, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            _pending_range.append((host, ev))
    return out


_pending_range = []


def _key_to_float(k):
    k = int(k) & 0xFFFFFFFF
    bits = (k & 0x7FFFFFFF) if (k & 0x80000000) else (~k & 0xFFFFFFFF)
    return float(np.array([bits], dtype=np.uint32).view(np.float32)[0])


def flush_range_warnings(block=False):
    """Print the reference's out-of-range warnings (data_utils.py:40-43) for every finished
    mel_spectrogram call; block=True waits for the outstanding ones."""
    keep = []
    for host, ev in _pending_range:
        if not block and not ev.query():
            keep.append((host, ev))
            continue
        ev.synchronize()
        lo, hi = -_key_to_float(host[0]), _key_to_float(host[1])
        if lo < -1.:
            print('min value is ', lo)
        if hi > 1.:
            print('max value is ', hi)
    _pending_range[:] = keep


# ---------------------------------------------------------------------------------------
# batching glue (pure views / cat; data_utils.py:158-178)
# ---------------------------------------------------------------------------------------
def combine_fixed_length(tensor_list, length):
    """Concatenate along time, zero-pad to a multiple of `length`, view as (n, length, ...)."""
    total_length = sum(t.size(0) for t in tensor_list)
    tensor_list = list(tensor_list)
    if total_length % length != 0:
        pad_length = length - (total_length % length)
        ref = tensor_list[0]
        tensor_list.append(torch.zeros(pad_length, *ref.size()[1:], dtype=ref.dtype,
                                       device=ref.device))
        total_length += pad_length
    tensor = torch.cat(tensor_list, 0)
    return tensor.view(total_length // length, length, *tensor.size()[1:])


def decollate_tensor(tensor, lengths):
    """Inverse of combine_fixed_length: slice the flattened (b*s, d) tensor by `lengths`."""
    b, s, d = tensor.size()
    flat = tensor.view(b * s, d)
    results = []
    idx = 0
    for length in lengths:
        assert idx + length <= b * s
        results.append(flat[idx:idx + length])
        idx += length
    return results


# ---------------------------------------------------------------------------------------
# host-side interface pieces used by EMGDataset consumers
# ---------------------------------------------------------------------------------------
class FeatureNormalizer(object):
    """data_utils.py:137-156; the attribute names are a pickle contract (normalizers.pkl)."""

    def __init__(self, feature_samples, share_scale=False):
        feature_samples = np.concatenate(feature_samples, axis=0)
        self.feature_means = feature_samples.mean(axis=0, keepdims=True)
        if share_scale:
            self.feature_stddevs = feature_samples.std()
        else:
            self.feature_stddevs = feature_samples.std(axis=0, keepdims=True)

    def normalize(self, sample):
        sample -= self.feature_means
        sample /= self.feature_stddevs
        return sample

    def inverse(self, sample):
        sample = sample * self.feature_stddevs
        sample = sample + self.feature_means
        return sample


class TextTransform(object):
    """data_utils.py:243-258 without the jiwer/unidecode dependencies: lower-case, strip
    punctuation, map to indices over a-z0-9 and space."""

    def __init__(self):
        self.chars = string.ascii_lowercase + string.digits + ' '

    def clean_text(self, text):
        try:
            from unidecode import unidecode
            text = unidecode(text)
        except ImportError:
            text = text.encode("ascii", "ignore").decode("ascii")
        text = text.translate(str.maketrans('', '', string.punctuation))
        return text.lower()

    def text_to_int(self, text):
        text = self.clean_text(text)
        return [self.chars.index(c) for c in text if c in self.chars]

    def int_to_text(self, ints):
        return ''.join(self.chars[i] for i in ints)


def _frame_rms_max(audio, frame_length=2048, hop_length=512):
    """max over frames of librosa.feature.rms(y=audio) (its defaults: centred frames of 2048,
    hop 512, zero padding), used by data_utils.py:19-27.  Host-side dataset preparation."""
    a = np.pad(np.asarray(audio, dtype=np.float64), frame_length // 2)
    n = 1 + (len(a) - frame_length) // hop_length
    if n <= 0:
        return float(np.sqrt(np.mean(a * a))) if len(a) else 0.0
    c = np.concatenate([[0.0], np.cumsum(a * a)])
    starts = np.arange(n) * hop_length
    power = (c[starts + frame_length] - c[starts]) / frame_length
    return float(np.sqrt(np.maximum(power, 0.0)).max())


def normalize_volume(audio):
    """data_utils.py:19-27."""
    max_rms = _frame_rms_max(audio) + 0.01
    audio = audio * (0.2 / max_rms)
    max_val = np.abs(audio).max()
    if max_val > 1.0:
        audio = audio / max_val
    return audio


def load_audio(filename, start=None, end=None, max_frames=None, renormalize_volume=False):
    """data_utils.py:64-83: read a clip (soundfile), optional volume normalisation and
    16 kHz -> 22.05 kHz resampling (host-side dataset preparation; resampling needs librosa like
    the reference), clip to [-1, 1], log-mel ON THE GPU (csrc/mel.cu).  Returns (frames, 80)
    numpy like the reference."""
    import soundfile as sf
    audio, r = sf.read(filename)
    if len(audio.shape) > 1:
        audio = audio[:, 0]
    if start is not None or end is not None:
        audio = audio[start:end]
    if renormalize_volume:
        audio = normalize_volume(audio)
    if r == 16000:
        try:
            import librosa
        except ImportError as e:
            raise _lib.SSBError(-3, "load_audio: 16 kHz input needs librosa.resample (host-side "
                                    "dataset preparation, not part of the GPU path)") from e
        audio = librosa.resample(audio, orig_sr=16000, target_sr=22050)
    else:
        assert r == 22050
    audio = np.clip(audio, -1, 1)
    y = torch.tensor(audio, dtype=torch.float32).unsqueeze(0).cuda()
    mspec = mel_spectrogram(y, 1024, 80, 22050, 256, 1024, 0, 8000, center=False)
    mspec = mspec.squeeze(0).T.cpu().numpy()
    flush_range_warnings(block=True)      # the copy above already synchronised: no extra wait
    if max_frames is not None and mspec.shape[0] > max_frames:
        mspec = mspec[:max_frames, :]
    return mspec

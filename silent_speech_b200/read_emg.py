"""Interface mirror of the reference's read_emg.py (EMGDataset :142-296, SizeAwareSampler
:115-140) backed by SYNTHETIC utterances.

The reference's dataset reads the Zenodo EMG corpus from disk and runs scipy signal filtering
(read_emg.py:27-100); that is CPU dataset preparation outside the hot path and
there is no corpus (or network) here.  What the hot path needs from this module is its
CONTRACT: the item dict schema, `collate_raw`, the sampler protocol and the attributes the
training scripts read.  This class honours that contract with seeded synthetic utterances of
the right shapes, dtypes and rate bookkeeping (8 raw EMG samples per 86.13 Hz feature frame).
"""
import random
from copy import copy

import numpy as np
import torch

from .data_utils import FeatureNormalizer, TextTransform


class _IdentityNormalizer(FeatureNormalizer):
    def __init__(self, dim):
        self.feature_means = np.zeros((1, dim), dtype=np.float32)
        self.feature_stddevs = np.ones((1, dim), dtype=np.float32)


def _maybe_pin(t):
    return t.pin_memory() if torch.cuda.is_available() else t


class EMGDataset(torch.utils.data.Dataset):
    """Same constructor arguments and public surface as read_emg.py:143-259; extra keyword
    arguments size the synthetic corpus."""

    def __init__(self, base_dir=None, limit_length=False, dev=False, test=False, no_testset=False,
                 no_normalizers=False, num_examples=64, frames=500, frames_jitter=0,
                 silent_fraction=0.5, seed=0):
        self.limit_length = limit_length
        self.no_normalizers = no_normalizers
        split = 1 if dev else (2 if test else 0)
        rs = random.Random(seed * 3 + split)
        self._items = []
        for i in range(num_examples):
            T = frames + (rs.randrange(-frames_jitter, frames_jitter + 1) if frames_jitter else 0)
            self._items.append({'frames': max(T, 8), 'silent': rs.random() < silent_fraction,
                                'seed': rs.randrange(1 << 30), 'session': i % 4})
        self.example_indices = list(range(num_examples))
        self.num_features = 112           # read_emg.py:197 (hand-crafted EMG features, unused by Model)
        self.num_speech_features = 80     # read_emg.py:196
        self.num_sessions = 4
        self.mfcc_norm = _IdentityNormalizer(80)
        self.emg_norm = _IdentityNormalizer(112)
        self.text_transform = TextTransform()
        self._cache = {}

    def silent_subset(self):
        result = copy(self)
        result.example_indices = [i for i in self.example_indices if self._items[i]['silent']]
        return result

    def subset(self, fraction):
        result = copy(self)
        result.example_indices = self.example_indices[:int(fraction * len(self.example_indices))]
        return result

    def __len__(self):
        return len(self.example_indices)

    def raw_length(self, i):
        """raw 1 kHz-equivalent length used by SizeAwareSampler's budget (read_emg.py:133)."""
        return self._items[self.example_indices[i]]['frames'] * 8

    def __getitem__(self, i):
        idx = self.example_indices[i]
        if idx in self._cache:
            return self._cache[idx]
        it = self._items[idx]
        g = torch.Generator().manual_seed(it['seed'])
        T = it['frames']
        Tg = int(1.2 * T) if it['silent'] else T          # SURVEY.md §8d target lengths
        raw = torch.randn(T * 8, 8, generator=g)           # O(1)-scale like 50*tanh(x/1000)
        text = "synthetic utterance %d" % idx
        result = {
            'audio_features': _maybe_pin(torch.randn(T, 80, generator=g)),
            'emg': _maybe_pin(torch.randn(T, 112, generator=g)),
            'text': text,
            'text_int': _maybe_pin(torch.tensor(self.text_transform.text_to_int(text),
                                                dtype=torch.int64)),
            'file_label': idx,
            'session_ids': _maybe_pin(torch.full((T,), it['session'], dtype=torch.int64)),
            'book_location': ('synthetic', idx),
            'silent': it['silent'],
            'raw_emg': _maybe_pin(raw),
        }
        if it['silent']:
            result['parallel_voiced_audio_features'] = _maybe_pin(torch.randn(Tg, 80, generator=g))
            result['parallel_voiced_emg'] = _maybe_pin(torch.randn(Tg, 112, generator=g))
        result['phonemes'] = _maybe_pin(torch.randint(0, 48, (Tg,), generator=g))
        result['audio_file'] = 'synthetic://%d' % idx
        self._cache[idx] = result
        return result

    @staticmethod
    def collate_raw(batch):
        """dict of lists, keys as read_emg.py:285-296."""
        audio_features, audio_feature_lengths, parallel_emg = [], [], []
        for ex in batch:
            if ex['silent']:
                audio_features.append(ex['parallel_voiced_audio_features'])
                audio_feature_lengths.append(ex['parallel_voiced_audio_features'].shape[0])
                parallel_emg.append(ex['parallel_voiced_emg'])
            else:
                audio_features.append(ex['audio_features'])
                audio_feature_lengths.append(ex['audio_features'].shape[0])
                parallel_emg.append(np.zeros(1))
        return {'audio_features': audio_features,
                'audio_feature_lengths': audio_feature_lengths,
                'emg': [ex['emg'] for ex in batch],
                'raw_emg': [ex['raw_emg'] for ex in batch],
                'parallel_voiced_emg': parallel_emg,
                'phonemes': [ex['phonemes'] for ex in batch],
                'session_ids': [ex['session_ids'] for ex in batch],
                'lengths': [ex['emg'].shape[0] for ex in batch],
                'silent': [ex['silent'] for ex in batch],
                'text_int': [ex['text_int'] for ex in batch],
                'text_int_lengths': [ex['text_int'].shape[0] for ex in batch]}


class SizeAwareSampler(torch.utils.data.Sampler):
    """read_emg.py:115-140: shuffled indices packed into batches of at most `max_len` raw
    samples; the last incomplete batch is dropped."""

    def __init__(self, emg_dataset, max_len):
        self.dataset = emg_dataset
        self.max_len = max_len

    def __iter__(self):
        indices = list(range(len(self.dataset)))
        random.shuffle(indices)
        batch, batch_length = [], 0
        for idx in indices:
            length = self.dataset.raw_length(idx)
            if length + batch_length > self.max_len:
                yield batch
                batch, batch_length = [], 0
            batch.append(idx)
            batch_length += length


def synthetic_batch(n_utt, frames, seed=1234, alternate_silent=True):
    """A collate_raw batch of `n_utt` utterances of `frames` feature frames (8*frames raw
    samples): silent / voiced alternate (16 silent + 16 voiced at bs=32)."""
    ds = EMGDataset(num_examples=n_utt, frames=frames, seed=seed)
    if alternate_silent:
        for i, it in enumerate(ds._items):
            it['silent'] = (i % 2 == 0)
    return EMGDataset.collate_raw([ds[i] for i in range(n_utt)])

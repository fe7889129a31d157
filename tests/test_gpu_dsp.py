"""DSP.wav_to_mel on the GPU against the numpy oracle (row a13)."""
import numpy as np
import pytest
import torch

from forwardtacotron_b200.utils import synth
from forwardtacotron_b200.utils.config import default_config
from forwardtacotron_b200.utils.dsp import DSP
from oracle import dsp_oracle

from util import GOLD, MAX_ABS, MEAN_ABS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dsp():
    return DSP.from_config(default_config())


def test_filterbank_matches_oracle(dsp):
    fb = dsp.mel_filterbank()
    want = dsp_oracle.mel_filterbank(22050, 1024, 80, 0, 8000)
    assert fb.shape == (80, 513) and np.abs(fb - want).max() < 1e-6
    assert int((fb != 0).sum()) == 727


def test_frozen_fixture(dsp):
    g = np.load(GOLD / 'dsp_noise10k.npz')
    mel = dsp.wav_to_mel(g['y'])
    assert isinstance(mel, np.ndarray) and mel.dtype == np.float32 and mel.shape == (80, 40)  # tests/test_dsp.py pins
    d = np.abs(mel - g['mel'])
    assert d.max() < MAX_ABS and d.mean() < MEAN_ABS, (d.max(), d.mean())
    print('dsp fixture max-abs', d.max(), 'mean-abs', d.mean())


@pytest.mark.parametrize('n', [1024, 1023, 513, 700, 10000, 44100, 220500, 255, 256, 257])
def test_lengths_and_edges(dsp, n):
    rng = np.random.default_rng(n)
    y = (0.1 * rng.standard_normal(n)).astype(np.float32)
    if n <= 512:
        pytest.skip('np.pad reflect needs len(y) > n_fft//2; librosa raises for such clips')
    want = dsp_oracle.wav_to_mel(y)
    got = dsp.wav_to_mel(y)
    assert got.shape == want.shape == (80, 1 + n // 256)
    d = np.abs(got - want)
    assert d.max() < MAX_ABS and d.mean() < MEAN_ABS, (n, d.max(), d.mean())


def test_silence_hits_the_clamp_floor(dsp):
    mel = dsp.wav_to_mel(np.zeros(5000, np.float32))
    assert np.allclose(mel, np.log(1e-5)) and np.isclose(mel.min(), -11.512925)
    lin = dsp.wav_to_mel(np.zeros(5000, np.float32), normalize=False)
    assert np.all(lin == 0)


def test_batch_of_ragged_clips_and_cuda_tensors(dsp):
    audio, offs = synth.synthetic_audio(12, seed=3, min_s=0.2, max_s=1.5)
    clips = [audio[int(offs[i]):int(offs[i + 1])].numpy() for i in range(12)]
    got = dsp.wav_to_mel_batch(clips)
    for i, c in enumerate(clips):
        want = dsp_oracle.wav_to_mel(c)
        assert got[i].shape == want.shape
        live = want > np.log(1e-5) + 2.0          # well above the clamp: fp32-FFT noise floor, see DESIGN.md
        d = np.abs(got[i] - want)
        worst = float(d[live].max()) if live.any() else 0.0   # a silent clip sits entirely on the clamp floor
        assert worst < MAX_ABS and d.mean() < MEAN_ABS, (i, worst, d.mean())
    # CUDA tensor in -> CUDA tensor out, same numbers
    t = dsp.wav_to_mel(torch.from_numpy(clips[0]).cuda())
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), got[0])
    packed, fo = dsp.wav_to_mel_packed(audio.cuda(), offs)
    assert packed.numel() == 80 * int(fo[-1])
    assert np.array_equal(packed[80 * int(fo[3]):80 * int(fo[4])].view(80, -1).cpu().numpy(), got[3])


def test_unsupported_fft_size_is_an_error():
    cfg = default_config()
    cfg['dsp']['n_fft'] = 2048
    with pytest.raises(Exception, match='1024'):
        DSP.from_config(cfg).wav_to_mel(np.zeros(4096, np.float32))


def test_featurize_folder_of_wavs(tmp_path):
    """preprocess.py's wav -> mel/{id}.npy step, batched on the GPU, against the numpy oracle (row f-2)."""
    from scipy.io import wavfile
    from forwardtacotron_b200 import preprocess
    from forwardtacotron_b200.utils.config import default_config
    from forwardtacotron_b200.utils.dsp import DSP
    from oracle import dsp_oracle
    rng = np.random.default_rng(3)
    wavs = {}
    for i, n in enumerate([22050, 30001, 4096]):
        y = (0.3 * rng.standard_normal(n)).astype(np.float32)
        if i == 1:
            y *= 5.0                                     # exceeds full scale -> peak scaling kicks in (preprocess.py:72)
        wavfile.write(str(tmp_path / f'clip{i}.wav'), 22050, y)
        wavs[f'clip{i}'] = y
    dsp = DSP.from_config(default_config())
    done = preprocess.featurize(sorted(tmp_path.glob('*.wav')), dsp, tmp_path / 'data')
    assert [d[0] for d in done] == ['clip0', 'clip1', 'clip2']
    for name, y in wavs.items():
        got = np.load(tmp_path / 'data' / 'mel' / f'{name}.npy')
        want = dsp_oracle.wav_to_mel(preprocess.peak_scale(y, False))
        assert got.dtype == np.float32 and got.shape == want.shape == (80, 1 + len(y) // 256)
        assert np.abs(got - want).max() < 1e-2 and np.abs(got - want).mean() < 1e-3

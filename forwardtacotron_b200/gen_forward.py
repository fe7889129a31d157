"""Batched counterpart of the reference's ``gen_forward.py`` (:43-134): checkpoint -> phonemised sentences -> mels,
written in the vocoder hand-off formats the reference produces (``.mel`` = torch.save for MelGAN :122, ``.npy`` for
HiFi-GAN :124).  WaveRNN / Griffin-Lim synthesis and the espeak cleaner are outside this package.

    python -m forwardtacotron_b200.gen_forward --checkpoint forward_step90k.pt --file sentences.phon.txt \\
        --alpha 1.0 --amp 1.0 --format npy [--exact]

The reference runs one sentence at a time.  Here sentences are bucketed by length and run as padded batches
(``utils/batching.py``).  ForwardTacotron batches carry every row's own length into the kernels
(``ForwardTacotron.generate_ragged``), so each sentence gets exactly the mel of its own one-sentence ``generate`` call
-- upstream's output, file for file, at batched throughput.  ``--no-mask`` runs the reference's no-mask arithmetic on
the padded batch instead (pad tokens are ordinary symbols there; rows are cut after their real tokens' frames);
``--exact`` forces one sentence per call (the only exact mode for FastPitch, which has no ragged path).
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import List

import numpy as np
import torch

from .utils import batching
from .utils.checkpoints import load_tts_model
from .utils.text import Tokenizer


def synthesize_texts(model, texts: List[str], alpha: float = 1.0, amp: float = 1.0, exact: bool = False,
                     max_tokens: int = 16384, no_mask: bool = False) -> List[torch.Tensor]:
    """phonemised strings -> list of (1, n_mels, L_i) CPU tensors (``gen['mel_post'].cpu()`` of the reference)."""
    tok = Tokenizer()
    utts = [tok(t) for t in texts]
    if any(len(u) == 0 for u in utts):
        raise ValueError('a sentence has no symbol of the phoneme inventory')
    pf = lambda p: p * amp      # gen_forward.py:103 "simple amplification of pitch"
    ef = lambda e: e            # gen_forward.py:104
    ragged = hasattr(model, 'generate_ragged') and not no_mask
    mels = batching.synthesize_corpus(model, utts, alpha=alpha, max_tokens=max_tokens, max_batch=1 if exact else 256,
                                      exact=ragged, pitch_function=pf, energy_function=ef)
    return [m.unsqueeze(0).cpu() for m in mels]


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description='Batched TTS mel generation on B200 (ForwardTacotron / FastPitch)')
    ap.add_argument('--checkpoint', required=True)
    ap.add_argument('--input_text', '-i', default=None, help='one phonemised sentence')
    ap.add_argument('--file', default='sentences.txt', help='phonemised sentences, one per line')
    ap.add_argument('--alpha', type=float, default=1.0, help='speed: durations are divided by alpha')
    ap.add_argument('--amp', type=float, default=1.0, help='pitch amplification')
    ap.add_argument('--format', choices=['mel', 'npy'], default='npy', help='.mel (MelGAN) or .npy (HiFi-GAN)')
    ap.add_argument('--exact', action='store_true', help='one sentence per call, as upstream')
    ap.add_argument('--no-mask', action='store_true', help="padded batches with the reference's no-mask arithmetic")
    ap.add_argument('--out', default='model_outputs')
    args = ap.parse_args(argv)

    model, config = load_tts_model(args.checkpoint)
    model = model.to('cuda')
    texts = [args.input_text] if args.input_text else \
        [l.strip() for l in Path(args.file).read_text(encoding='utf-8').splitlines() if l.strip()]
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    k = model.get_step() // 1000
    mels = synthesize_texts(model, texts, args.alpha, args.amp, args.exact, no_mask=args.no_mask)
    for i, m in enumerate(mels, 1):
        name = f'{i}_forward_{k}k_alpha{args.alpha}_amp{args.amp}_{"melgan" if args.format == "mel" else "hifigan"}'
        if args.format == 'mel':
            torch.save(m, out / f'{name}.mel')
        else:
            np.save(out / f'{name}.npy', m.numpy(), allow_pickle=False)
    print(f'wrote {len(mels)} mels to {out}/')


if __name__ == '__main__':
    main()

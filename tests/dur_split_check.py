#!/usr/bin/env python
"""Accuracy of the duration predictor paths against an fp64 evaluation of the reference algorithm:
split-precision tensor-core GEMMs (default) vs the fp32 SIMT GEMM (FTB_OPT_DUR_SIMT) vs the fp32 CPU oracle.
    python tests/dur_split_check.py        (needs a B200)"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from forwardtacotron_b200 import _lib  # noqa: E402
from forwardtacotron_b200.utils import synth  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402


def rounded(d):
    return (d.float() + 0.5).long()


def truth64(sd64, x):
    p = 'dur_pred'
    v = sd64[p + '.embedding.weight'][x].transpose(1, 2)
    for i in range(3):
        v = mo.conv_relu_bn(sd64, f'{p}.convs.{i}', v, relu=True)
    v = mo.rnn_explicit(sd64, p + '.rnn', v.transpose(1, 2), 'gru')
    return (v @ sd64[p + '.lin.weight'].T + sd64[p + '.lin.bias']).squeeze(-1)


def main():
    model, _ = synth.synthetic_model('forward_tacotron')
    sd32 = {k: v.clone() for k, v in model.state_dict().items()}
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd32.items()}
    model = model.cuda()
    lib = _lib.lib()
    for B, T, seed in ((64, 200, 1), (64, 200, 2), (64, 200, 3), (128, 300, 5), (16, 900, 7)):
        x = synth.synthetic_tokens(B, T, seed=seed)
        truth = truth64(sd64, x)
        cpu32 = mo.ft_series_predictor(sd32, 'dur_pred', x).squeeze(-1)
        res = {}
        for name, simt in (('split', 0), ('simt', 1)):
            h = model._get_handle(torch.device('cuda', 0))
            _lib.check(lib.ftb_ft_set_option(h, _lib.FTB_OPT_DUR_SIMT, simt))
            d = model.run_series_predictor('dur_pred', x.cuda()).squeeze(-1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                d = model.run_series_predictor('dur_pred', x.cuda()).squeeze(-1)
            torch.cuda.synchronize()
            res[name] = (d.cpu(), (time.perf_counter() - t0) / 5 * 1e3)
        _lib.check(lib.ftb_ft_set_option(h, _lib.FTB_OPT_DUR_SIMT, 0))
        line = f'B{B} T{T} seed{seed}: dur mean {float(truth.mean()):.2f}'
        for name, d in (('split', res['split'][0]), ('simt', res['simt'][0]), ('cpu32', cpu32)):
            err = (d.double() - truth).abs()
            flips = int((rounded(d) != rounded(truth)).sum())
            line += f' | {name}: max {float(err.max()):.2e} mean {float(err.mean()):.2e} flips {flips}'
        line += f' | ms split {res["split"][1]:.3f} simt {res["simt"][1]:.3f}'
        line += f' | split==simt rounded: {bool(torch.equal(rounded(res["split"][0]), rounded(res["simt"][0])))}'
        print(line, flush=True)
    print('tc timeouts', lib.ftb_tc_timeout_count())


if __name__ == '__main__':
    main()

"""In-tree build of the sm_100a extension: every csrc/*.cu -> one shared library
``forwardtacotron_b200/csrc/libftb200.so`` (C ABI in include/ftb200.h).

nvcc cross-compiles without a GPU.  Object files are cached on the source
mtime; the .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / 'csrc'
LIB = CSRC / 'libftb200.so'
OBJ_DIR = CSRC / 'build'
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC']
# developer builds: FTB_NVCC_DEFINES="FTB_PHASE_TIMING" compiles the clock stamps of the scripts/*_phase_timing.py tools in
# (they are compiled OUT by default: a clock64() read is a scheduling barrier and cost the GRU-256 step 8 %)
NVCC_FLAGS += ['-D' + d for d in os.environ.get('FTB_NVCC_DEFINES', '').split() if d]


def _nvcc() -> str:
    cand = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(cand):
        raise RuntimeError('nvcc not found: the ftb200 extension cannot be built')
    return cand


def _newest_header() -> float:
    hdrs = list(CSRC.glob('*.cuh')) + [CSRC.parent.parent / 'include' / 'ftb200.h']
    return max(h.stat().st_mtime for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = sorted(CSRC.glob('*.cu'))
    OBJ_DIR.mkdir(exist_ok=True)
    hdr_m = _newest_header()
    nvcc = _nvcc()
    jobs = []
    for src in srcs:
        obj = OBJ_DIR / (src.stem + '.o')
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, '-c', str(src), '-o', str(obj)]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}')
        return src.name, r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name, log in ex.map(compile_one, jobs):
                if verbose and log:
                    print(f'--- {name}\n{log}', file=sys.stderr)
    objs = [str(OBJ_DIR / (s.stem + '.o')) for s in srcs]
    if jobs or not LIB.exists():
        # static cudart: the library only needs libcuda at run time (resolved through cudart)
        cmd = [nvcc, '-shared', '-o', str(LIB), *objs, '-gencode', 'arch=compute_100a,code=sm_100a',
               '-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

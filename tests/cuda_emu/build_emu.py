"""TEST INFRASTRUCTURE ONLY: compiles the product's .cu sources with g++ against the CPU
emulator (cuda_emu.h) into tests/cuda_emu/_build/libstb_emu.so.  See cuda_emu.h."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
CSRC = os.path.join(ROOT, 'scannertools_b200', 'csrc')
OUT = os.path.join(HERE, '_build', 'libstb_emu.so')
SOURCES = ['common.cu', 'hist.cu', 'farneback.cu', 'pipe.cu', 'resize.cu', 'convert_color.cu', 'hist_hsv.cu']


def build(force=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(HERE, 'cuda_emu.h'), os.path.join(HERE, 'cuda_emu.cpp'),
                   os.path.join(CSRC, 'stb_rt.h'), os.path.join(ROOT, 'include', 'stb.h')]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ['g++', '-O2', '-g', '-std=c++17', '-fPIC', '-shared', '-pthread', '-DSTB_CPU_EMU_BUILD',
           '-ffp-contract=off', '-Wno-unknown-pragmas',
           '-I', HERE, '-I', CSRC, '-I', os.path.join(ROOT, 'include'), '-o', OUT]
    for s in srcs:
        cmd += ['-x', 'c++', s]
    cmd += ['-x', 'c++', os.path.join(HERE, 'cuda_emu.cpp')]
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force=True))

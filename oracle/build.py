"""Build the oracle's C restatement (TEST INFRASTRUCTURE): gcc -> oracle/_build/liboracle_nms.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'liboracle_nms.so')


def build(force=False):
    src = os.path.join(HERE, 'nms_greedy.c')
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-o', LIB, src])
    return LIB


if __name__ == '__main__':
    print(build(force=True))

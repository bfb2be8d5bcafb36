"""The two-pass register FFT of the log-mel kernel (csrc/logmel_fft.cuh) is __host__ __device__: its lane-level passes are
compiled for the CPU here and run lane by lane in the kernel's phase order against a float64 DFT (index math, twiddle
tables, bit reversal, the untangle pairs and the transposed power layout).  No GPU and nothing from oracle/ involved."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_pass_fft_lane_math_on_cpu(tmp_path):
    cxx = shutil.which("g++")
    inc = "/usr/local/cuda/include"
    if cxx is None or not os.path.isfile(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path / "lfft_host")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-I", inc, "-I", os.path.join(ROOT, "whisperseg_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "logmel_fft_host.cpp"), "-o", exe])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    worst = float(res.stdout.strip().splitlines()[-1].split()[1])
    assert worst < 2e-6            # relative to the frame's largest bin; fp32 FFT of 512 / 1024 points

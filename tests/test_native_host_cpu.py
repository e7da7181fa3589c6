"""examples/native_host.cpp (a host with neither Python nor torch) compiles against include/abcnet_b200.h + the shared library alone,
and without a CUDA device it stops with a message and a non-zero status instead of computing anything on the CPU. The run on a
B200 is tests/test_native_gpu.py::test_native_host_program."""
import os
import struct
import subprocess

import torch

from oracle import synth, unet_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_native_host_compiles_and_refuses_to_run_without_a_device(tmp_path):
    exe = tmp_path / "native_host"
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", os.path.join(ROOT, "examples", "native_host.cpp"), "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(cuda, "include"), "-L", os.path.join(ROOT, "abcnet_b200"), "-labcnet_b200",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{os.path.join(ROOT, 'abcnet_b200')}", "-o", str(exe)],
                   check=True)
    if torch.cuda.is_available():
        return                                              # the GPU suite runs it for real
    sd = unet_ref.make_state_dict(seed=31, variant="W1")
    with open(tmp_path / "weights.bin", "wb") as f:
        for k, v in sd.items():
            if torch.is_floating_point(v):
                name = k.encode()
                f.write(struct.pack("<i", len(name)) + name + struct.pack("<q", v.numel()) + v.detach().float().contiguous().numpy().tobytes())
    x = torch.from_numpy(synth.binary_images(31, 1, 64, 64, 0.08))
    (tmp_path / "images.u8").write_bytes((x > 0).to(torch.uint8).numpy().tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "weights.bin"), str(tmp_path / "images.u8"), "1", "64", "64"], capture_output=True, text=True,
                       env=dict(os.environ, LD_LIBRARY_PATH=os.path.join(cuda, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", "")))
    assert r.returncode != 0 and r.stdout.strip() == "" and "failed" in r.stderr, (r.returncode, r.stdout, r.stderr)

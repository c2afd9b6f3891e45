"""The C ABI from C: tests/c/cabi_consumer.c is a plain C99 program that links libb200va.so (what a cgo / JNI /
N-API shim or a C++ host would do).  CPU: the header compiles as C and every symbol the program uses resolves
at link time.  GPU: the program's letterbox / detections / tracks equal the oracle's, bit for bit."""
import os
import shutil
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(REPO, "realtime_video_analytics_32streams_b200", "lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_consumer(out_dir):
    exe = os.path.join(out_dir, "cabi_consumer")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", f"-I{os.path.join(REPO, 'include')}",
           f"-I{os.path.join(CUDA, 'include')}", os.path.join(REPO, "tests", "c", "cabi_consumer.c"), "-o", exe,
           f"-L{LIB_DIR}", "-lb200va", f"-L{os.path.join(CUDA, 'lib64')}", "-lcudart", f"-Wl,-rpath,{LIB_DIR}",
           f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_c_consumer_compiles_and_links(tmp_path):
    from realtime_video_analytics_32streams_b200.build import build

    build()
    exe = build_consumer(str(tmp_path))
    ldd = subprocess.run(["ldd", exe], stdout=subprocess.PIPE, text=True).stdout
    assert "libb200va.so" in ldd and "not found" not in ldd.split("libb200va.so")[1].splitlines()[0]


@pytest.mark.gpu
def test_c_consumer_matches_oracle(tmp_path):
    from oracle import hotpath as O
    from realtime_video_analytics_32streams_b200 import synth

    exe = build_consumer(str(tmp_path))
    H, W, C, A = 1080, 1920, 84, 8400
    frame = synth.synth_frame(4242, H, W)
    head = synth.synth_head(4243, C, A, 14, dup=3)
    frame.tofile(tmp_path / "frame.bin")
    np.ascontiguousarray(head, dtype=np.float32).tofile(tmp_path / "head.bin")
    proc = subprocess.run([exe, str(tmp_path / "frame.bin"), str(H), str(W), str(tmp_path / "head.bin"), str(C), str(A),
                           str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout
    ref, meta = O.preprocess(frame, (640, 640))
    net = np.fromfile(tmp_path / "net.bin", dtype=np.float32).reshape(3, 640, 640)
    assert np.array_equal(net.view(np.uint32), ref[0].view(np.uint32))
    dets = O.filter_detections(O.postprocess(head[None], meta, 0.35, 0.5), 0.35)
    assert f"dets={len(dets)} " in proc.stdout and len(dets) > 0
    assert np.fromfile(tmp_path / "det_cls.bin", dtype=np.int32).tolist() == [d.class_id for d in dets]
    assert np.array_equal(np.fromfile(tmp_path / "det_conf.bin", dtype=np.float32),
                          np.array([d.confidence for d in dets], dtype=np.float32))
    assert np.array_equal(np.fromfile(tmp_path / "det_box.bin", dtype=np.float32).reshape(-1, 4),
                          np.array([d.bbox_xyxy for d in dets], dtype=np.float32))
    trk = O.IouTracker(30, 0.5, 1)
    trk.update("s", dets)
    want = trk.update("s", dets)
    assert np.fromfile(tmp_path / "trk_id.bin", dtype=np.int64).tolist() == [t.track_id for t in want]
    assert np.fromfile(tmp_path / "trk_hits.bin", dtype=np.int32).tolist() == [t.hits for t in want]
    assert np.array_equal(np.fromfile(tmp_path / "trk_box.bin", dtype=np.float64).reshape(-1, 4),
                          np.array([t.bbox_xyxy for t in want], dtype=np.float64))

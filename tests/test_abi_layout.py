"""The ctypes mirrors in _native.py against the C structs of include/b200va.h: a C program compiled with gcc
prints sizeof / offsetof for every struct that crosses the boundary, and the numbers must equal what ctypes
lays out.  Catches silent drift between the header and the Python binding (no GPU needed)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STRUCTS = {
    "b200va_config": ("Config", ["device", "max_batch", "max_anchors", "max_candidates", "max_dets", "max_streams", "max_tracks"]),
    "b200va_letterbox": ("Letterbox", ["src_h", "src_w", "new_h", "new_w", "pad_left", "pad_top", "scale"]),
    "b200va_dets": ("Dets", ["bbox_xyxy", "conf", "cls", "count"]),
    "b200va_dets64": ("Dets64", ["bbox_xyxy", "conf", "cls", "count"]),
    "b200va_tracker_cfg": ("TrackerCfg", ["max_age", "min_hits", "max_iou_distance"]),
    "b200va_tracks": ("Tracks", ["track_id", "cls", "conf", "bbox_xyxy", "age", "hits", "count", "rows"]),
    "b200va_tick_args": ("TickArgs", None),  # every field, in ctypes order
}


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_ctypes_structs_match_the_header(tmp_path):
    from realtime_video_analytics_32streams_b200 import _native as N

    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "b200va.h"', "int main(void) {"]
    expected = {}
    for cname, (pyname, fields) in STRUCTS.items():
        cls = getattr(N, pyname)
        if fields is None:
            fields = [f[0] for f in cls._fields_]
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        expected[cname] = C.sizeof(cls)
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
            expected[f"{cname}.{f}"] = getattr(cls, f).offset
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    proc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{os.path.join(REPO, 'include')}", str(src), "-o", str(exe)],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout
    got = {k: int(v) for k, v in (line.split() for line in out.strip().splitlines())}
    assert got == expected


def test_enum_values_match_the_header():
    """The integer constants the binding passes are the header's enum values."""
    import re
    from realtime_video_analytics_32streams_b200 import _native as N

    text = open(os.path.join(REPO, "include", "b200va.h")).read()
    want = {"B200VA_OUT_F32_RGB_NCHW": N.OUT_F32_RGB_NCHW, "B200VA_OUT_F16_RGB_NCHW": N.OUT_F16_RGB_NCHW,
            "B200VA_OUT_U8_BGR_NCHW": N.OUT_U8_BGR_NCHW, "B200VA_OUT_U8_BGR_NHWC": N.OUT_U8_BGR_NHWC,
            "B200VA_HEAD_CHANNEL_MAJOR": N.HEAD_CHANNEL_MAJOR, "B200VA_HEAD_ANCHOR_MAJOR": N.HEAD_ANCHOR_MAJOR,
            "B200VA_SCORE_REF_COMPAT": N.SCORE_REF_COMPAT, "B200VA_SCORE_V8_NATIVE": N.SCORE_V8_NATIVE,
            "B200VA_NMS_AGNOSTIC": N.NMS_AGNOSTIC, "B200VA_NMS_CLASS_AWARE": N.NMS_CLASS_AWARE,
            "B200VA_OK": N.OK, "B200VA_ERR_INVALID": N.ERR_INVALID, "B200VA_ERR_CUDA": N.ERR_CUDA,
            "B200VA_ERR_CAPACITY": N.ERR_CAPACITY, "B200VA_ERR_STATE": N.ERR_STATE}
    for name, value in want.items():
        m = re.search(rf"\b{name}\s*=\s*(-?\d+)", text)
        assert m and int(m.group(1)) == value, name
    m = re.search(r"#define\s+B200VA_OUT_FLAG_PADS_VALID\s+(0x[0-9a-fA-F]+)", text)
    assert m and int(m.group(1), 16) == N.OUT_FLAG_PADS_VALID

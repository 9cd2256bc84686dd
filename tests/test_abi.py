"""The C-ABI library loads without a GPU and exports every symbol include/erd_b200.h declares;
argument checking works on the host.  No compute call is made here."""
import ctypes as C
import os
import re

from erd_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'erd_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(erd_[a-z_0-9]+)\s*\(', text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = N.load()
    names = _declared_functions()
    assert len(names) >= 15
    for nm in names:
        assert hasattr(lib, nm), f'{nm} declared in erd_b200.h but not exported'
        assert nm in N.SIGNATURES, f'{nm} has no ctypes signature in erd_b200/_native.py'
    assert lib.erd_abi_version() == N.ABI_VERSION


def _shape(n=16, levels=((100, 168), (50, 84), (25, 42), (13, 21), (7, 11)), C_=80, ori=40):
    s = N.ErdShape()
    s.num_imgs, s.num_levels, s.num_classes, s.ori_classes, s.reg_max = n, len(levels), C_, ori, 16
    for i, (h, w) in enumerate(levels):
        s.level_h[i], s.level_w[i], s.stride[i] = h, w, 8 << i
    s.anchor_scale, s.kd_temperature = 8.0, 10.0
    s.loss_weight_cls, s.loss_weight_bbox, s.loss_weight_dfl, s.loss_weight_ld = 1.0, 2.0, 0.25, 0.25
    return s


def test_sizes_for_the_baseline_config():
    lib, z = N.load(), N.ErdSizes()
    assert lib.erd_sizes(C.byref(_shape()), C.byref(z)) == 0
    assert z.anchors_per_img == 22400 and z.sel_cap == 22400 // 5 + 1 and z.num_losses == 15 + 32
    assert z.workspace_bytes > 16 * 22400 * 28


def test_bad_shapes_are_rejected_with_status_codes():
    lib, z = N.load(), N.ErdSizes()
    s = _shape()
    s.num_levels = 4
    assert lib.erd_sizes(C.byref(s), C.byref(z)) == -1 and b'num_levels' in lib.erd_last_error()
    s = _shape()
    s.reg_max = 8
    assert lib.erd_sizes(C.byref(s), C.byref(z)) == -1
    s = _shape(ori=80)
    assert lib.erd_sizes(C.byref(s), C.byref(z)) == -1
    assert lib.erd_sizes(None, C.byref(z)) == -2
    assert lib.erd_sizes(C.byref(_shape()), None) == -2


def test_null_pointers_are_rejected_before_any_launch():
    lib = N.load()
    s = _shape()
    nul = N.PtrArray()
    assert lib.erd_ers_select(C.byref(s), nul, nul, None, None, None, None, None, None, None, None) == -2
    assert lib.erd_atss_assign(C.byref(s), None, None, None, None, None, None, None, None) == -2
    assert lib.erd_teacher_nms(C.byref(s), None, None, None, 0.005, None, None, None, None, None) == -2
    assert lib.erd_launch_count() == 0


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The header is plain C: compile a probe with gcc and compare sizeof / offsetof of every
    struct that crosses the boundary with the ctypes mirror in erd_b200/_native.py."""
    import os
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        import pytest
        pytest.skip('no gcc')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {'ErdShape': N.ErdShape, 'ErdSizes': N.ErdSizes, 'ErdStepBuffers': N.ErdStepBuffers,
               'ErdPredictConfig': N.ErdPredictConfig, 'ErdTeacherHead': N.ErdTeacherHead}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "erd_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'probe.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'probe'
    subprocess.run([gcc, '-std=c99', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(out[f'{name}.{field}']) == getattr(cls, field).offset, f'{name}.{field}'

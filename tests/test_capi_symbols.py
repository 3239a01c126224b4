"""The C-ABI library loads without a GPU and exports every symbol include/qlb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "qlb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qlb200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_functions():
    names = declared_functions()
    assert "qlb200_match_create" in names and "qlb200_execute" in names and len(names) >= 35


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "tensortoolkit_b200", "libqlb200.so"))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_header():
    from tensortoolkit_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_functions()
    assert _lib.lib.qlb200_version().startswith(b"qlb200")


def test_no_cpu_fallback_without_device():
    """On a box without a usable sm_100 GPU the execution entry points must fail loudly."""
    import tensortoolkit_b200 as tk
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(tk._lib.QLB200Error):
        tk.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensortoolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "refbridge", "contract_np", "libqlref", '#include "oracle'):
                    assert bad not in src, f"{f} uses the oracle ({bad})"
    for f in ("qlb200.h", os.path.join("qlten_b200", "contract.h")):
        assert "oracle/" not in open(os.path.join(ROOT, "include", f)).read()


def test_header_is_plain_c_and_the_partitioner_answers_from_c(tmp_path):
    """include/qlb200.h must compile as C (the boundary is a C ABI: plain pointers and sizes), and a C program linked
    against the library gets the known answer of a small partition: two sectors of 100 rows, the second twice as costly per
    row, two ranks -> the cut lies at row 25 of the second sector (150 of 300 weight units), snapped to a multiple of 8."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "cuts.c"
    src.write_text(r'''
#include <stdio.h>
#include "qlb200.h"
int main(void) {
  qlb200_piece line[2] = {{0, 0, 100, 0, 1.0}, {1, 0, 100, 0, 2.0}};
  uint32_t degs[2] = {100, 100}, ranges[2][2][2];
  if (qlb200_shard_cut_line(line, 2, degs, 2, 2, 8, &ranges[0][0][0]) != QLB200_OK) return 2;
  printf("%u %u %u %u | %u %u %u %u\n", ranges[0][0][0], ranges[0][0][1], ranges[0][1][0], ranges[0][1][1],
         ranges[1][0][0], ranges[1][0][1], ranges[1][1][0], ranges[1][1][1]);
  double t[2] = {1.0, 3.0};
  qlb200_piece out[4];
  unsigned long long n = qlb200_shard_reweigh(line, 2, &ranges[0][0][0], 2, 2, t, 1.0, 4, out);
  printf("%llu %s\n", n, qlb200_version());
  return 0;
}
''')
    exe = tmp_path / "cuts"
    lib_dir = os.path.join(root, "tensortoolkit_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-lqlb200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "0 100 0 24 | 0 0 24 100"          # 25 snaps to 24
    assert out[1].split()[0] == "3"                      # sector 0 (rank 0), sector 1 split at the cut

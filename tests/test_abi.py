"""The C-ABI library loads on a CPU-only box and exports every symbol include/asmgpu.h declares;
compute entry points refuse to run without a device (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from asm_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(name="asmgpu.h", prefix="asm_"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z0-9_]+)\s*\(" % prefix, src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 35
    raw = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"libasmgpu.so does not export {n}"
    # the binding covers the header one to one
    assert sorted(L.SIGNATURES) == names
    assert L.lib().asm_version().startswith(b"asm-b200")
    # host wrapper (criteria mirror): include/asmgpu_host.h
    from asm_b200 import criteria
    host = header_functions("asmgpu_host.h", "asmh_")
    assert len(host) >= 20 and sorted(criteria.HOST_SIGNATURES) == host
    for n in host:
        assert hasattr(raw, n), f"libasmgpu.so does not export {n}"


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "asmgpu.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)  # declarations only, comments stripped
    assert "torch" not in code and "at::" not in code and "std::" not in code and "#include <stdint.h>" in code


@pytest.mark.skipif(L.lib().asm_device_count() > 0, reason="CPU-only behaviour")
def test_no_cpu_fallback():
    with pytest.raises(L.AsmGpuError) as e:
        L.GpuContext(2)
    assert e.value.code == L.ASM_E_NO_DEVICE and "no CPU fallback" in str(e.value)
    # null context: every entry point returns ASM_E_INVALID instead of crashing
    assert L.lib().asm_psrf_eval(None, None, 0, 0, 1, 0, None, None) == L.ASM_E_INVALID
    assert L.lib().asm_ess_batch(None, 1, 1, 1, None, 2000, None, None) == L.ASM_E_INVALID
    from asm_b200 import criteria
    with pytest.raises(L.AsmGpuError) as e:
        criteria.TraceInfo(2)  # TraceInfo owns a GPU context
    assert e.value.code == L.ASM_E_NO_DEVICE


def test_criteria_validation_matches_the_reference():
    """initAndValidate sites (TreePSRF.java:54-84; TreeESS.java:45-76) -- host logic, no GPU needed."""
    from asm_b200 import criteria as K
    K.TreePSRF()  # defaults are valid
    K.TreePSRF(smoothing=0.0), K.TreePSRF(smoothing=0.9, cacheLimit=200)
    for kw in (dict(smoothing=1.5), dict(smoothing=-0.1), dict(smoothing=float("nan")), dict(targetESS=0),
               dict(cacheLimit=0), dict(cacheLimit=100, targetESS=100), dict(sampleSize=0),
               dict(sampleSize=100, targetESS=100)):
        with pytest.raises(K.IllegalArgumentException):
            K.TreePSRF(**kw)
    # TreeESS.initAndValidate as written throws for every smoothing >= 0, its own default included (TreeESS.java:47)
    with pytest.raises(K.IllegalArgumentException):
        K.TreeESS()
    K.TreeESS(smoothing=-0.5)
    assert K.TraceESS().header() == "TraceESS" and K.TreePSRF().header() == "TreePSRF"


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "asm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"


def test_headers_are_plain_c(tmp_path):
    """include/asmgpu.h and include/asmgpu_host.h compile as C11 (what cgo / JNI / Panama jextract would consume) and a
    C program calling a no-GPU entry point links against the library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "asmgpu.h"\n#include "asmgpu_host.h"\n#include <stdio.h>\n'
        'int main(void) { asm_shard s = {0, 1}; (void)s; printf("%s %d\\n", asm_version(), (int)asm_device_count());\n'
        '  int64_t lines = -1; const char* p[1] = {"/nonexistent"}; int32_t b[1] = {0};\n'
        '  int32_t rc = asmh_combine_logs(0, 1, p, b, 1, 1, 0, 0, "/dev/null", &lines);\n'
        '  return rc == ASM_OK ? 1 : 0; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = tmp_path / "abi"
    libdir = os.path.dirname(L.LIB_PATH)
    r = subprocess.run([gcc, "-std=c11", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-l:libasmgpu.so",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("asm-b200")  # the missing chain file is an error, not a crash


def test_java_stub_handles_match_the_header():
    """java/ cannot be compiled here (no JDK): at least every Panama downcall handle names an exported function and
    passes as many arguments, of the right width, as include/asmgpu.h declares."""
    src = open(os.path.join(ROOT, "java", "asm", "inference", "AsmGpuNative.java")).read()
    handles = re.findall(r'h\("(asm_[a-z0-9_]+)",\s*FunctionDescriptor\.of\(([^;]*?)\)\);', src, flags=re.S)
    assert len(handles) >= 10
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "asmgpu.h")).read(), flags=re.S)
    width = {"JAVA_INT": ("int32_t",), "JAVA_LONG": ("int64_t",), "ADDRESS": ("*",), "JAVA_DOUBLE": ("double",)}
    for name, desc in handles:
        m = re.search(r"\b(\w[\w\s\*]*?)\b%s\s*\(([^)]*)\)\s*;" % name, hdr)
        assert m, f"{name} is not declared in asmgpu.h"
        c_args = [a.strip() for a in m.group(2).split(",")] if m.group(2).strip() not in ("", "void") else []
        j = [a.strip() for a in desc.split(",")]
        j_ret, j_args = j[0], j[1:]
        assert len(j_args) == len(c_args), f"{name}: {len(j_args)} Java arguments, {len(c_args)} in the header"
        for ja, ca in zip(j_args, c_args):
            if ja == "ADDRESS":
                assert "*" in ca, f"{name}: ADDRESS for `{ca}`"
            else:
                assert "*" not in ca and any(w in ca for w in width[ja]), f"{name}: {ja} for `{ca}`"
        assert ("*" in m.group(1)) == (j_ret == "ADDRESS")


def test_bind_thread_near_device_is_harmless():
    """asm_bind_thread_near_device: an error code without a device, otherwise the number of CPUs bound to (0 where the
    platform shows one NUMA node or hides the topology); the calling thread keeps a non-empty CPU set either way."""
    before = os.sched_getaffinity(0)
    rc = L.lib().asm_bind_thread_near_device(0)
    after = os.sched_getaffinity(0)
    if L.lib().asm_device_count() == 0:
        assert rc == L.ASM_E_INVALID and after == before
    else:
        assert rc >= 0 and len(after) > 0 and (rc == 0 or len(after) == rc)
    assert L.lib().asm_bind_thread_near_device(10 ** 6) == L.ASM_E_INVALID
    os.sched_setaffinity(0, before)

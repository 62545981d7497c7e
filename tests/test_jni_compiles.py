"""The JNI shim (java/src/main/native/needle_jni.c) must keep compiling.  There is no JDK in the build image, so it is
compiled against a compile-only stub of <jni.h> (java/src/test/native/stub/jni.h) with warnings as errors, and every
native method NeedleNative.java declares must have its Java_..._name definition.  CPU only."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "java", "src", "main", "native", "needle_jni.c")


def test_jni_shim_compiles_against_the_stub_header(tmp_path):
    obj = tmp_path / "needle_jni.o"
    cmd = ["gcc", "-std=c11", "-fPIC", "-Wall", "-Wextra", "-Werror", "-c", SHIM, "-o", str(obj),
           "-I", os.path.join(ROOT, "java", "src", "test", "native", "stub"), "-I", os.path.join(ROOT, "include")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    syms = subprocess.run(["nm", "--defined-only", str(obj)], capture_output=True, text=True).stdout
    with open(os.path.join(ROOT, "java", "src", "main", "java", "com", "justinblank", "strings", "gpu", "NeedleNative.java")) as f:
        natives = re.findall(r"static native [\w.\[\]]+ (\w+)\(", f.read())
    assert len(natives) >= 9
    for name in natives:
        assert f"Java_com_justinblank_strings_gpu_NeedleNative_{name}" in syms, name


def test_jni_shim_holds_no_critical_region():
    with open(SHIM) as f:
        code = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    assert "Critical" not in code  # blocking library calls must not sit inside Get*Critical (JNI spec; stalls the collector)

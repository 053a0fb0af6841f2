"""The JNI shim (jni/pinot_b200_jni.c) against the C ABI: compiled with the repository's minimal stand-in for <jni.h>
(jni/stub/jni.h -- there is no JDK in the build image) with -Wall -Werror, and its exported Java_* symbols must be exactly the
native methods org.apache.pinot.b200.Native declares."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jni_shim_compiles_against_the_c_abi_and_matches_native_java():
    with tempfile.TemporaryDirectory() as d:
        obj = os.path.join(d, "pinot_b200_jni.o")
        subprocess.check_call(["gcc", "-std=gnu11", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror", "-fPIC", "-DPB_WITH_JNI",
                               "-I" + os.path.join(ROOT, "jni", "stub"), "-I" + os.path.join(ROOT, "include"),
                               "-c", os.path.join(ROOT, "jni", "pinot_b200_jni.c"), "-o", obj])
        nm = subprocess.check_output(["nm", "-g", "--defined-only", obj]).decode()
    exported = {m.group(1) for m in re.finditer(r" T Java_org_apache_pinot_b200_Native_(\w+)", nm)}
    java = open(os.path.join(ROOT, "java", "org", "apache", "pinot", "b200", "Native.java")).read()
    declared = set(re.findall(r"static native [\w\[\]]+ (\w+)\(", java))
    assert exported == declared, (sorted(exported - declared), sorted(declared - exported))
    # every C entry point the shim calls is declared in the public header
    shim = open(os.path.join(ROOT, "jni", "pinot_b200_jni.c")).read()
    header = open(os.path.join(ROOT, "include", "pinot_b200.h")).read()
    for fn in set(re.findall(r"\b(pb_[a-z_0-9]+)\(", shim)):
        assert re.search(r"\b" + fn + r"\(", header), fn

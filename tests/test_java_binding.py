"""The Java FFM binding cannot be compiled here (no JDK in the image or on the GPU box), so its downcall descriptors are
checked against the ctypes table the tests DO exercise: same entry points, same argument count, same argument classes
(pointer / 64-bit / 32-bit / double).  A descriptor that drifts from include/vw_modwt.h fails here, not in a JVM."""
import ctypes as C
import os
import re

from vectorwave_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java", "com", "morphiqlabs", "wavelet")


def _klass(t):
    if t is None:
        return "VOID"
    if t in (C.c_double,):
        return "JAVA_DOUBLE"
    if t in (C.c_int64, C.c_uint64, C.c_size_t):
        return "JAVA_LONG"
    if t in (C.c_int, C.c_int32, C.c_uint32):
        return "JAVA_INT"
    return "ADDRESS"      # c_void_p, c_char_p, POINTER(...)


def _descriptors():
    text = open(os.path.join(JAVA, "gpu", "VwNative.java")).read()
    out = {}
    for m in re.finditer(r'h\("(vw_[a-z0-9_]+)",\s*FunctionDescriptor\.(of|ofVoid)\(([^;]*?)\)\);', text, re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",") if a.strip()]
        ret = "VOID" if m.group(2) == "ofVoid" else args.pop(0)
        out[m.group(1)] = (ret, args)
    return out


def test_every_java_downcall_matches_the_ctypes_signature():
    desc = _descriptors()
    assert len(desc) >= 40
    for name, (ret, args) in desc.items():
        assert name in _native.SIGNATURES, f"{name}: bound in Java but not declared in the ABI table"
        res, argtypes = _native.SIGNATURES[name]
        assert ret == _klass(res), f"{name}: return {ret} vs {_klass(res)}"
        assert args == [_klass(t) for t in argtypes], f"{name}: {args} vs {[_klass(t) for t in argtypes]}"


def test_java_facades_only_use_bound_handles_and_existing_classes():
    bound = set(_descriptors())
    java_files = []
    for d, _, files in os.walk(os.path.join(ROOT, "java")):
        java_files += [os.path.join(d, f) for f in files if f.endswith(".java")]
    assert len(java_files) >= 6
    classes = {os.path.basename(f)[:-5] for f in java_files}
    for f in java_files:
        text = open(f).read()
        for m in re.finditer(r"VwNative\.(vw_[a-z0-9_]+)\.invokeExact", text):
            assert m.group(1) in bound, f"{os.path.basename(f)} calls unbound handle {m.group(1)}"
    # INTEGRATION.md must only name Java classes that exist under java/
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in re.findall(r"\b(Gpu[A-Za-z]+|PinnedArena|VwNative)\b", doc):
        assert name in classes, f"INTEGRATION.md names {name}, which does not exist under java/"


def test_span_plan_struct_size_matches_java_constant():
    text = open(os.path.join(JAVA, "gpu", "VwNative.java")).read()
    m = re.search(r"SPAN_PLAN_BYTES = ([^;]+);", text)
    assert eval(m.group(1)) == C.sizeof(_native.VwSpanPlan)

"""CPU: the C-ABI shared library loads and exports every function include/mtb200.h declares (no compute calls)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_if_needed():
    so = os.path.join(ROOT, "mangatranslator_b200", "lib", "libmtb200.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    return so


def test_library_exports_every_declared_symbol():
    so = _build_if_needed()
    lib = C.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "mtb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mtb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 8
    for n in sorted(names):
        assert hasattr(lib, n), f"libmtb200.so does not export {n}"


def test_struct_layouts_match_header_sizes():
    from mangatranslator_b200 import clean_host as H
    # sizes as the C compiler lays the header structs out (natural alignment)
    assert C.sizeof(H.CleanParams) % 8 == 0
    assert C.sizeof(H.CleanJob) % 8 == 0
    assert C.sizeof(H.CleanResult) % 8 == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import importlib
    import mangatranslator_b200._lib as L
    monkeypatch.setattr(L, "LIB_PATH", tmp_path / "nope.so")
    monkeypatch.setattr(L, "_lib", None)
    try:
        L.lib()
        raised = False
    except L.MtbError:
        raised = True
    assert raised
    importlib.reload(L)


def test_header_is_plain_c_and_ctypes_structs_match_the_compiler(tmp_path):
    """include/mtb200.h is the drop-in boundary: it must compile as C99 (no C++ / torch types), and every ctypes mirror
    in the Python layer must have exactly the size the C compiler gives the header's struct."""
    import json
    import subprocess
    from mangatranslator_b200 import _lib, clean_host, conjoined, safebox_host, sam2, yolo
    hdr = os.path.join(ROOT, "include", "mtb200.h")
    subprocess.check_call(["gcc", "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", hdr])
    pairs = {"mtb_conv_desc": _lib.ConvDesc, "mtb_clean_params": clean_host.CleanParams, "mtb_clean_job": clean_host.CleanJob,
             "mtb_clean_result": clean_host.CleanResult, "mtb_yolo_level": yolo.YoloLevel, "mtb_nms_params": yolo.NmsParams,
             "mtb_split_pair": conjoined.SplitPair, "mtb_attn_desc": sam2.AttnDesc,
             "mtb_safebox_job": safebox_host.SafeBoxJob, "mtb_safebox_result": safebox_host.SafeBoxResult}
    src = tmp_path / "sizes.c"
    body = "".join(f'  printf("%s\\"{n}\\": %zu", first ? "" : ", ", sizeof({n})); first = 0;\n' for n in pairs)
    src.write_text(f'#include <stdio.h>\n#include "{hdr}"\nint main(void) {{ int first = 1; printf("{{");\n{body}  printf("}}\\n"); return 0; }}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe), str(src)])
    sizes = json.loads(subprocess.check_output([str(exe)]).decode())
    for name, cls in pairs.items():
        assert C.sizeof(cls) == sizes[name], (name, C.sizeof(cls), sizes[name])
    # ... and every field sits at the compiler's offset (the ctypes mirrors use the header's field names)
    lines = "".join(f'  printf("{n}.{f[0]} %zu\\n", offsetof({n}, {f[0]}));\n' for n, cls in pairs.items() for f in cls._fields_)
    src2 = tmp_path / "offsets.c"
    src2.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{hdr}"\nint main(void) {{\n{lines}  return 0; }}\n')
    exe2 = tmp_path / "offsets"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe2), str(src2)])
    checked = 0
    for line in subprocess.check_output([str(exe2)]).decode().splitlines():
        key, off = line.split()
        n, f = key.split(".")
        assert getattr(pairs[n], f).offset == int(off), key
        checked += 1
    assert checked > 100

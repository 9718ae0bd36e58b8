"""The Java FFM shim (integration/java) cannot be compiled here (no JDK in the image); what can be checked is that it
cannot drift from the C ABI: NptLayouts.java is regenerated from include/npt.h with gcc and compared, every symbol the
binding looks up is declared in the header and exported by libnpt.so with the same number of arguments, and every layout
constant the sources use exists."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JDIR = os.path.join(ROOT, "integration", "java", "npTranscript", "cluster")
sys.path.insert(0, os.path.join(ROOT, "integration"))


def _read(name):
    return open(os.path.join(JDIR, name)).read()


def test_layouts_are_generated_from_the_header():
    import gen_layouts
    assert gen_layouts.render() == _read("NptLayouts.java"), "run python integration/gen_layouts.py"


def test_every_downcall_names_an_exported_symbol_with_matching_arity():
    from nptranscript_b200 import abi, engine
    L = engine.lib()
    src = _read("Npt.java")
    calls = re.findall(r'h\("(npt_\w+)",\s*FunctionDescriptor\.(ofVoid|of)\(([^;]*?)\)\);', src, re.S)
    assert len(calls) >= 15
    header = open(os.path.join(ROOT, "include", "npt.h")).read()
    for name, kind, args in calls:
        assert name in abi.SYMBOLS and hasattr(L, name), name
        n_args = len([a for a in args.split(",") if a.strip()]) - (1 if kind == "of" else 0)
        decl = re.search(r"\b" + name + r"\s*\(([^)]*)\)\s*;", header)
        assert decl, name
        params = [p for p in decl.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert n_args == len(params), (name, n_args, params)


def test_layout_constants_used_by_the_sources_exist():
    layouts = set(re.findall(r"static final (?:long|int) (\w+) =", _read("NptLayouts.java")))
    for f in os.listdir(JDIR):
        if f.endswith(".java") and f != "NptLayouts.java":
            for c in re.findall(r"NptLayouts\.(\w+)", _read(f)):
                assert c in layouts, (f, c)

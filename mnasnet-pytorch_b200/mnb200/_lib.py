"""ctypes binding of libmnb200.so.  Signatures are derived from include/mnb200.h so the header stays the
single source of truth for the C ABI.  There is NO fallback: if the library is missing and cannot be
built, importing this module raises."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HEADER = os.path.join(ROOT, "include", "mnb200.h")
LIB_PATH = os.path.join(HERE, "libmnb200.so")

MNB_F32, MNB_BF16 = 0, 1
LAYOUT_NHWC, LAYOUT_NCHW_F32, LAYOUT_NHWC_U8 = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2

_SCALARS = {
    "int": ctypes.c_int, "long long": ctypes.c_longlong, "unsigned long long": ctypes.c_ulonglong,
    "float": ctypes.c_float, "double": ctypes.c_double,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [(argname, ctype)])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    decls = {}
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(mnb_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argl.append((a.split("*")[-1].strip(), ctypes.c_void_p))
                else:
                    toks = a.replace("const ", "").split()
                    ty, nm = " ".join(toks[:-1]), toks[-1]
                    argl.append((nm, _SCALARS[ty]))
        decls[name] = (restype, argl)
    return decls


DECLS = parse_header()


def _load():
    from . import build
    stale = False
    try:
        stale = build.needs_build()          # sources / header newer than the library
    except Exception:
        stale = not os.path.exists(LIB_PATH)
    if stale or os.environ.get("MNB200_REBUILD"):
        try:
            build.build(force=bool(os.environ.get("MNB200_REBUILD")))
        except Exception:
            if not os.path.exists(LIB_PATH):
                raise                        # no library and no way to build it: fail loudly
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argl) in DECLS.items():
        fn = getattr(lib, name)          # AttributeError -> header/library mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = [t for _, t in argl]
    return lib


lib = _load()


class MnbError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib.mnb_last_error()
        raise MnbError(f"{what} failed rc={rc}: {msg.decode() if msg else ''}")


def call(name, *args):
    check(getattr(lib, name)(*args), name)


OPTION_EPOCH = 0        # bumped by set_option: captured CUDA graphs of a launch program are only valid within one epoch


def set_option(name: str, value: int):
    """Process-wide kernel-selection switch ("pw_stream", "stem_mma", "dw_mma", "dw_small", ...; include/mnb200.h)."""
    global OPTION_EPOCH
    check(lib.mnb_set_option(name.encode(), int(value)), "set_option")
    OPTION_EPOCH += 1


def get_option(name: str) -> int:
    v = lib.mnb_get_option(name.encode())
    if v < 0:
        check(v, "get_option")
    return v


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()

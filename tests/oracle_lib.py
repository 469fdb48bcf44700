"""ctypes binding of the CPU oracle (oracle/nrgrep_oracle.c) -- test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
import ctypes
import os
import subprocess
ENV = dict(os.environ, GLIBC_TUNABLES='glibc.malloc.tcache_count=0:glibc.malloc.perturb=255')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libnrgrep_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "nrgrep_coords")

NRO_MAXM = 256
NRO_MAXK = 15
TYPE_NAMES = {0: "SIMPLE", 1: "SPLIT", 2: "BWD", 3: "FWD"}


class NroPattern(ctypes.Structure):
    _fields_ = [("m", ctypes.c_int), ("start_line", ctypes.c_int), ("end_line", ctypes.c_int),
                ("cls", (ctypes.c_uint64 * 4) * NRO_MAXM)]


class NroPlan(ctypes.Structure):
    _fields_ = [("k", ctypes.c_int), ("ins", ctypes.c_int), ("del_", ctypes.c_int), ("subs", ctypes.c_int),
                ("type", ctypes.c_int), ("L", ctypes.c_int), ("npieces", ctypes.c_int),
                ("V", ctypes.c_int * (NRO_MAXK + 1)),
                ("split_cost", ctypes.c_double), ("fb_cost", ctypes.c_double),
                ("fb_flag", ctypes.c_int), ("fb_beg", ctypes.c_int), ("fb_end", ctypes.c_int)]


class NrxPattern(ctypes.Structure):          # EXTENDED patterns (?, *, +), k = 0
    _fields_ = [("P", NroPattern), ("op", ctypes.c_ubyte * NRO_MAXM)]


class NrxPlan(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("anchor", ctypes.c_int), ("wlen", ctypes.c_int), ("beg", ctypes.c_int),
                ("end", ctypes.c_int), ("cost", ctypes.c_double)]


class NroHit(ctypes.Structure):
    _fields_ = [("beg", ctypes.c_int64), ("end", ctypes.c_int64)]


_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "_build/libnrgrep_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = ctypes.CDLL(LIB_PATH)
        L.nro_parse.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(NroPattern)]
        L.nro_parse.restype = ctypes.c_int
        L.nro_plan_make.argtypes = [ctypes.POINTER(NroPattern), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.POINTER(NroPlan)]
        L.nro_plan_make.restype = ctypes.c_int
        L.nro_search.argtypes = [ctypes.POINTER(NroPattern), ctypes.POINTER(NroPlan), ctypes.c_char_p,
                                 ctypes.c_int64, ctypes.POINTER(NroHit), ctypes.c_int64]
        L.nro_search.restype = ctypes.c_int64
        L.nro_search_buffered.argtypes = [ctypes.POINTER(NroPattern), ctypes.POINTER(NroPlan), ctypes.c_char_p,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(NroHit), ctypes.c_int64]
        L.nro_search_buffered.restype = ctypes.c_int64
        L.nrx_parse.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(NrxPattern)]
        L.nrx_parse.restype = ctypes.c_int
        L.nrx_plan_make.argtypes = [ctypes.POINTER(NrxPattern), ctypes.POINTER(NrxPlan)]
        L.nrx_plan_make.restype = ctypes.c_int
        L.nrx_search_buffered.argtypes = [ctypes.POINTER(NrxPattern), ctypes.POINTER(NrxPlan), ctypes.c_char_p,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(NroHit), ctypes.c_int64]
        L.nrx_search_buffered.restype = ctypes.c_int64
        _lib = L
    return _lib


def parse_kopt(kopt):
    """'2ids' -> (2, ins, del, subs) exactly as main @400e00 parses -k."""
    digits = kopt.rstrip("idst")
    flags = kopt[len(digits):]
    k = int(digits) if digits else 0
    if not flags:
        return k, 1, 1, 1
    return k, int("i" in flags), int("d" in flags), int("s" in flags)


def plan(pattern, kopt="0ids", icase=True):
    L = lib()
    P = NroPattern()
    rc = L.nro_parse(pattern.encode("latin-1"), int(icase), ctypes.byref(P))
    if rc:
        raise ValueError("oracle cannot parse %r: %d" % (pattern, rc))
    k, i, d, s = parse_kopt(kopt)
    pl = NroPlan()
    rc = L.nro_plan_make(ctypes.byref(P), k, i, d, s, ctypes.byref(pl))
    if rc:
        raise ValueError("oracle cannot plan %r: %d" % (pattern, rc))
    return P, pl


def set_compat(deployed):
    """piece choice of the deployed binary (default allocator) instead of the defined zero-scratch behaviour"""
    lib().nro_set_compat(1 if deployed else 0)


BUFSIZE = 1600000          # patmatch.py:37 MAX_BUFFER_SIZE, passed as -b


def plan_ext(pattern, icase=True):
    """EXTENDED pattern (k = 0): -> (NrxPattern, NrxPlan); ValueError when the oracle does not cover it."""
    L = lib()
    X = NrxPattern()
    rc = L.nrx_parse(pattern.encode("latin-1"), int(icase), ctypes.byref(X))
    if rc == 1:                                   # NRX_REWRITTEN_SIMPLE: the parser's rewrites left a SIMPLE pattern in X.P
        return X, None
    if rc:
        raise ValueError("oracle cannot parse %r as an EXTENDED pattern: %d" % (pattern, rc))
    pl = NrxPlan()
    rc = L.nrx_plan_make(ctypes.byref(X), ctypes.byref(pl))
    if rc:
        raise ValueError("oracle does not cover the plan of %r: %d" % (pattern, rc))
    return X, pl


def is_extended(pattern):
    return any(ch in pattern for ch in "?*+")


def search(pattern, text, kopt="0ids", icase=True, cap=1 << 20, bufsize=BUFSIZE):
    """-> list of (beg, end) byte offsets, the hit list `nrgrep_coords -b bufsize` prints."""
    L = lib()
    if isinstance(text, str):
        text = text.encode("latin-1")
    hits = (NroHit * cap)()
    if is_extended(pattern):
        if parse_kopt(kopt)[0] != 0:
            raise ValueError("EXTENDED patterns with errors are not covered by the oracle")
        X, xpl = plan_ext(pattern, icase)
        if xpl is None:
            spl = NroPlan()
            L.nro_plan_make(ctypes.byref(X.P), 0, 1, 1, 1, ctypes.byref(spl))
            n = L.nro_search_buffered(ctypes.byref(X.P), ctypes.byref(spl), text, len(text), bufsize, hits, cap)
        else:
            n = L.nrx_search_buffered(ctypes.byref(X), ctypes.byref(xpl), text, len(text), bufsize, hits, cap)
        if n > cap:
            raise OverflowError("oracle hit buffer too small: %d" % n)
        return [(hits[i].beg, hits[i].end) for i in range(n)]
    P, pl = plan(pattern, kopt, icase)
    n = L.nro_search_buffered(ctypes.byref(P), ctypes.byref(pl), text, len(text), bufsize, hits, cap)
    if n > cap:
        raise OverflowError("oracle hit buffer too small: %d" % n)
    return [(hits[i].beg, hits[i].end) for i in range(n)]


def run_reference(pattern, path, kopt="0ids", binary=None, bufsize=BUFSIZE):
    """Run the reference engine itself exactly as patmatch.py:733 does; -> (banner, [(beg, end, text)])."""
    import re
    binary = binary or REF_BIN
    out = subprocess.run([binary, "-i", "-b", str(bufsize), "-k", kopt, pattern, path],
                         capture_output=True, env=ENV).stdout.decode("latin-1")
    hits = []
    lines = out.split("\n")
    for l in lines:
        m = re.match(r"\[(\d+), (\d+)\]: (.*)$", l)
        if m:
            hits.append((int(m.group(1)), int(m.group(2)), m.group(3)))
    return (lines[0] if lines else ""), hits

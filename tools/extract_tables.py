#!/usr/bin/env python3
"""Build-time table extractor.

Reads the numeric literal tables of the reference decoder (read-only, from
/root/reference) and emits them as *data* into one generated C header that the
oracle, the generator and the CUDA engine all include.  Tables are data, not
code: several of them (FFT twiddles, KBD windows, SBR/PS prototypes) are short
decimal literals that are NOT the correctly rounded values of their defining
formulas, so parity requires the very same float32 bit patterns
(SURVEY.md §7.3, Appendix B).

Float literals are converted the way javac does it: decimal string -> nearest
binary32, round-half-even, in ONE rounding step (never via double).  Floats are
emitted as uint32 bit patterns so no C compiler can re-round them.

Usage:  python tools/extract_tables.py [--ref /root/reference] [--out <header>]
        python tools/extract_tables.py --check     (exit 1 if header is stale)
"""
from __future__ import annotations

import argparse
import os
import re
import struct
import sys
from fractions import Fraction

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_REF = "/root/reference"
DEFAULT_OUT = os.path.join(ROOT, "jaadec_b200", "csrc", "generated", "jaad_tables.h")
AAC = "aac/src/main/java/net/sourceforge/jaad/aac/"


# --------------------------------------------------------------------------
# decimal literal -> binary32, single correctly-rounded step (javac semantics)
# --------------------------------------------------------------------------
def _f32_bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def _bits_f32(b: int) -> float:
    return struct.unpack("<f", struct.pack("<I", b & 0xFFFFFFFF))[0]


def dec_to_f32_bits(tok: str) -> int:
    """Correctly rounded decimal -> float32 bit pattern."""
    t = tok.rstrip("fFdD")
    exact = Fraction(t)
    neg = exact < 0 or t.strip().startswith("-")
    mag = abs(exact)
    if mag == 0:
        return 0x80000000 if neg else 0
    # candidate from the double path, then repair against its two neighbours
    try:
        c = _f32_bits(abs(float(t)))
    except OverflowError:
        c = 0x7F7FFFFF
    best = None
    for cand in (c - 1, c, c + 1):
        if cand < 0 or cand > 0x7F800000:
            continue
        if cand == 0x7F800000:
            continue
        err = abs(Fraction(_bits_f32(cand)) - mag)
        key = (err, cand & 1)  # ties -> even mantissa
        if best is None or key < best[0]:
            best = (key, cand)
    bits = best[1]
    return bits | (0x80000000 if neg else 0)


# --------------------------------------------------------------------------
# Java initializer parser
# --------------------------------------------------------------------------
_NUM = re.compile(r"[-+]?(?:0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)[fFdDlL]?")


def _strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    return src


def _find_initializer(src: str, name: str) -> str:
    m = re.search(r"\b" + re.escape(name) + r"\s*(?:\[\s*\]\s*)*=\s*(?:new\s+\w+\s*(?:\[\s*\]\s*)+)?\{", src)
    if not m:
        raise KeyError(name)
    i = m.end() - 1
    depth = 0
    for j in range(i, len(src)):
        ch = src[j]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                return src[i : j + 1]
    raise ValueError("unbalanced initializer for " + name)


def _parse_nested(text: str, conv):
    """Parse '{a, b, {c, d}}' into nested python lists using conv(token)."""
    pos = 0

    def parse():
        nonlocal pos
        assert text[pos] == "{"
        pos += 1
        out = []
        while True:
            while pos < len(text) and text[pos] in " \t\r\n,":
                pos += 1
            if text[pos] == "}":
                pos += 1
                return out
            if text[pos] == "{":
                out.append(parse())
                continue
            m = _NUM.match(text, pos)
            if not m:
                # identifiers (e.g. references to other tables) are returned verbatim
                m2 = re.compile(r"[A-Za-z_][A-Za-z_0-9.]*").match(text, pos)
                if not m2:
                    raise ValueError("cannot parse at %r" % text[pos : pos + 40])
                out.append(m2.group(0))
                pos = m2.end()
                continue
            out.append(conv(m.group(0)))
            pos = m.end()

    return parse()


def _int_tok(tok: str) -> int:
    t = tok.rstrip("lL")
    return int(t, 0)


class JavaFile:
    def __init__(self, ref: str, rel: str):
        with open(os.path.join(ref, rel), "r", encoding="utf-8", errors="replace") as f:
            self.src = _strip_comments(f.read())

    def floats(self, name):
        return _parse_nested(_find_initializer(self.src, name), dec_to_f32_bits)

    def ints(self, name):
        return _parse_nested(_find_initializer(self.src, name), _int_tok)


# --------------------------------------------------------------------------
# emit helpers
# --------------------------------------------------------------------------
def _flatten(x):
    if isinstance(x, list):
        for e in x:
            yield from _flatten(e)
    else:
        yield x


def _shape(x):
    s = []
    while isinstance(x, list):
        s.append(len(x))
        x = x[0]
    return s


class Emitter:
    def __init__(self):
        self.lines = []

    def raw(self, s=""):
        self.lines.append(s)

    def f32(self, cname, nested, comment=""):
        """float table stored as uint32 bit patterns (+ a float view accessor)."""
        shp = _shape(nested)
        flat = list(_flatten(nested))
        n = 1
        for d in shp:
            n *= d
        assert n == len(flat), (cname, shp, len(flat))
        dims = "".join("[%d]" % d for d in shp)
        self.raw("/* %s  shape %s  %s */" % (cname, "x".join(map(str, shp)), comment))
        self.raw("JAAD_TABLE_F32(%s, %d, \"%s\")" % (cname, n, dims))
        self._body(["0x%08Xu" % v for v in flat], 8)

    def i32(self, cname, nested, ctype="int32_t", comment=""):
        shp = _shape(nested)
        flat = list(_flatten(nested))
        self.raw("/* %s  shape %s  %s */" % (cname, "x".join(map(str, shp)), comment))
        self.raw("JAAD_TABLE_INT(%s, %s, %d)" % (cname, ctype, len(flat)))
        self._body([str(v) for v in flat], 16)

    def _body(self, toks, per):
        for i in range(0, len(toks), per):
            self.raw("  " + ", ".join(toks[i : i + per]) + ("," if i + per < len(toks) else ""))
        self.raw("JAAD_TABLE_END")
        self.raw()


HEADER = """\
/* GENERATED by tools/extract_tables.py -- do not edit.
 *
 * Numeric tables (data) of the JAAD reference decoder, as float32 bit
 * patterns / integers.  Every table names the reference file it was read
 * from.  Include with JAAD_TABLE_F32 / JAAD_TABLE_INT / JAAD_TABLE_END
 * defined by the includer (see jaad_tables_host.h for the default).
 */
"""


def build(ref: str) -> str:
    em = Emitter()
    em.raw(HEADER)

    # ---- Huffman codebooks: rows {len, codeword, values...} ---------------
    cb = JavaFile(ref, AAC + "huffman/Codebooks.java")
    for k in range(1, 12):
        rows = cb.ints("HCB%d" % k)
        width = 6 if k < 5 else 4
        for r in rows:
            assert len(r) == width, (k, r)
        em.i32("HCB%d" % k, rows, "int32_t", "huffman/Codebooks.java rows of {len,code,%d values}" % (width - 2))
    rows = cb.ints("HCB_SF")
    em.i32("HCB_SF", rows, "int32_t", "huffman/Codebooks.java rows of {len,code,value}")

    # ---- noiseless tables --------------------------------------------------
    em.f32("SCALEFACTOR_TABLE", JavaFile(ref, AAC + "syntax/ScaleFactorTable.java").floats("SCALEFACTOR_TABLE"),
           "syntax/ScaleFactorTable.java:9")
    em.f32("IQ_TABLE", JavaFile(ref, AAC + "syntax/IQTable.java").floats("IQ_TABLE"), "syntax/IQTable.java:11")

    sfb = JavaFile(ref, AAC + "syntax/ScaleFactorBands.java")
    em.i32("SWB_LONG_WINDOW_COUNT", sfb.ints("SWB_LONG_WINDOW_COUNT"), "int32_t", "syntax/ScaleFactorBands.java:7")
    em.i32("SWB_SHORT_WINDOW_COUNT", sfb.ints("SWB_SHORT_WINDOW_COUNT"), "int32_t", "syntax/ScaleFactorBands.java:71")
    long_names = sfb.ints("SWB_OFFSET_LONG_WINDOW")
    short_names = sfb.ints("SWB_OFFSET_SHORT_WINDOW")
    long_tab, short_tab = [], []
    for nm in long_names:
        t = sfb.ints(nm)
        assert len(t) <= 53
        long_tab.append(t + [-1] * (53 - len(t)))
    for nm in short_names:
        t = sfb.ints(nm)
        assert len(t) <= 17
        short_tab.append(t + [-1] * (17 - len(t)))
    em.i32("SWB_OFFSET_LONG", long_tab, "int16_t", "per sf_index, padded to 53 with -1")
    em.i32("SWB_OFFSET_SHORT", short_tab, "int16_t", "per sf_index, padded to 17 with -1")

    tns = JavaFile(ref, AAC + "tools/TNSTables.java")
    for nm in ("TNS_COEF_0_3", "TNS_COEF_0_4", "TNS_COEF_1_3", "TNS_COEF_1_4"):
        em.f32(nm, tns.floats(nm), "tools/TNSTables.java")

    # ---- filterbank ----------------------------------------------------------
    md = JavaFile(ref, AAC + "filterbank/MDCTTables.java")
    em.f32("MDCT_TABLE_2048", md.floats("MDCT_TABLE_2048"), "filterbank/MDCTTables.java:5")
    em.f32("MDCT_TABLE_128", md.floats("MDCT_TABLE_128"), "filterbank/MDCTTables.java:519")
    ff = JavaFile(ref, AAC + "filterbank/FFTTables.java")
    em.f32("FFT_TABLE_512", ff.floats("FFT_TABLE_512"), "filterbank/FFTTables.java:5  {re, im_inverse, im_forward}")
    em.f32("FFT_TABLE_64", ff.floats("FFT_TABLE_64"), "filterbank/FFTTables.java:519 {re, im_inverse}")
    sw = JavaFile(ref, AAC + "filterbank/SineWindows.java")
    em.f32("SINE_1024", sw.floats("SINE_1024"), "filterbank/SineWindows.java:5")
    em.f32("SINE_128", sw.floats("SINE_128"), "filterbank/SineWindows.java:1031")
    kw = JavaFile(ref, AAC + "filterbank/KBDWindows.java")
    em.f32("KBD_1024", kw.floats("KBD_1024"), "filterbank/KBDWindows.java:5")
    em.f32("KBD_128", kw.floats("KBD_128"), "filterbank/KBDWindows.java:1031")

    extra = os.path.join(os.path.dirname(os.path.abspath(__file__)), "extract_tables_sbr.py")
    if os.path.exists(extra):
        import importlib.util

        spec = importlib.util.spec_from_file_location("extract_tables_sbr", extra)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.emit(em, ref, JavaFile, AAC)

    return "\n".join(em.lines) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=DEFAULT_REF)
    ap.add_argument("--out", default=DEFAULT_OUT)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    text = build(a.ref)
    if a.check:
        with open(a.out) as f:
            ok = f.read() == text
        print("tables header is %s" % ("up to date" if ok else "STALE"))
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        f.write(text)
    print("wrote %s (%d bytes)" % (a.out, len(text)))


if __name__ == "__main__":
    main()

"""Seeded synthetic inputs for BASELINE.json's configs; generator in csrc/synth.c.

Both arms of bench.py (the CUDA path and the CPU reference arm) and the parity tests consume the FASTA
texts produced here, so they see byte-identical inputs.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsynth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            from . import build
            build.build()
        lib = C.CDLL(_SO)
        lib.synth_family.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_char_p, C.c_char_p, C.c_size_t]
        lib.synth_family.restype = C.c_size_t
        lib.synth_read.argtypes = [C.c_uint64, C.c_char_p, C.c_uint32, C.c_uint32, C.c_double, C.c_char_p, C.c_size_t]
        lib.synth_read.restype = C.c_size_t
        _lib = lib
    return _lib


def family(seed, n_seqs, root_len, rate, prefix="f"):
    """One family FASTA: members are independent mutations of one random root (star phylogeny)."""
    cap = n_seqs * (2 * root_len + 64) + 64
    buf = C.create_string_buffer(cap)
    n = _load().synth_family(seed, n_seqs, root_len, rate, prefix.encode(), buf, cap)
    assert n > 0
    return buf.raw[:n]


def config2_family(f, n_seqs=20, root_len=1000, rate=0.05):
    """Config 2: gene families, 20 seqs x 1 kb, 5% edits (seed 0xC2_0000_0000 + f)."""
    return family(0xC200000000 + f, n_seqs, root_len, rate, prefix=f"f{f}")


def config3_window(w, n_reads=30, template_len=500, rate=0.10):
    """Config 3: consensus windows, 30 reads x 500 bp, 10% edits (seed 0xC3_0000_0000 + w)."""
    return family(0xC300000000 + w, n_reads, template_len, rate, prefix=f"w{w}")


def config4_family(n_seqs=1000, root_len=1500, rate=0.15):
    """Config 4: one family of 1000 16S-like 1.5 kb sequences at 15% divergence."""
    return family(0xC400000000, n_seqs, root_len, rate, prefix="c4")


def config5(n_haps=50, root_len=20000, hap_rate=0.01, n_reads=1000, read_len=2000, read_rate=0.05):
    """Config 5: haplotype FASTA for the graph and reads sampled uniformly from the haplotypes."""
    hap_fa = family(0xC500000000, n_haps, root_len, hap_rate, prefix="hap")
    haps = [s for _, s in parse_fasta(hap_fa)]
    lib = _load()
    reads = []
    buf = C.create_string_buffer(2 * read_len + 16)
    for r in range(n_reads):
        h = haps[(r * 2654435761) % n_haps]
        n = lib.synth_read(0xC500000001 + r, h, len(h), read_len, read_rate, buf, len(buf))
        reads.append(b">read%d\n" % r + buf.raw[:n] + b"\n")
    return hap_fa, b"".join(reads)


def config5_reads(n_haps, root_len, hap_rate, first, n_reads, read_len=2000, read_rate=0.05):
    """Config 5 with an explicit read range: (haplotype FASTA, [read first .. first + n_reads) as byte strings)."""
    hap_fa = family(0xC500000000, n_haps, root_len, hap_rate, prefix="hap")
    haps = [s for _, s in parse_fasta(hap_fa)]
    lib = _load()
    buf = C.create_string_buffer(2 * read_len + 16)
    reads = []
    for r in range(first, first + n_reads):
        h = haps[(r * 2654435761) % n_haps]
        n = lib.synth_read(0xC500000001 + r, h, len(h), read_len, read_rate, buf, len(buf))
        reads.append(buf.raw[:n])
    return hap_fa, reads


def parse_fasta(text):
    recs = []
    for block in text.split(b">")[1:]:
        lines = block.split(b"\n")
        recs.append((lines[0].split()[0].decode(), b"".join(l.strip() for l in lines[1:])))
    return recs

"""Synthetic sequence files of the shapes BASELINE.json names.

Every base is a pure function of (seed, global base index) -- a counter-based generator -- so
the numpy implementation here and the CUDA generator in ``csrc/synth.cu`` produce bit-identical
files of any size without ever materialising one on the other side:

    base(g)      = "ACGT"[ splitmix64(seed + g) >> 62 ]
    N-run        : one run per N_SLOT bases, length from a 16-quantile table of an exponential
                   with mean ~1 kbp, start uniform inside the slot            (config 5)
    soft-mask run: one lower-case run per LC_SLOT bases, mean ~300 bp         (config 5)

File layout (what ``k6thru11fullANDupstream.sh`` feeds the reference):
    record r = header line (fixed width, ends in '\\n') + its bases in lines of ``line_width``
    columns, every line '\\n'-terminated (``line_width == 0``: one unwrapped line, the
    ``get_upstreams.pl:86-90`` record shape).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MASK64 = (1 << 64) - 1
N_SLOT = 20000
LC_SLOT = 3000
# 16 quantiles of an exponential distribution with mean 1000 (mean of the table = 955)
RUN_QUANTILES = np.array([32, 98, 170, 247, 330, 421, 520, 629, 752, 891, 1052, 1242, 1474, 1773, 2197, 3466],
                         dtype=np.uint64)
N_SALT = 0x9E3779B97F4A7C15
LC_SALT = 0xC2B2AE3D27D4EB4F


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser over uint64 (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


@dataclass(frozen=True)
class FastaLayout:
    """Byte layout of a synthetic multi-record file."""
    n_bases: int
    n_records: int
    line_width: int          # > 0 wrapped; 0 = one unwrapped line per record; < 0 = stripped layout (no newlines)
    header_fmt: str          # must render to the same width for every record index
    seed: int
    n_runs: bool = False     # config 5: N runs
    soft_mask: bool = False  # config 5: lower-case runs

    def header(self, r: int) -> bytes:
        return (self.header_fmt % r).encode("ascii") if "%" in self.header_fmt else self.header_fmt.encode("ascii")

    @property
    def header_len(self) -> int:
        return len(self.header(0))

    def record_bases(self, r: int) -> int:
        per = self.n_bases // self.n_records
        return per if r < self.n_records - 1 else self.n_bases - per * (self.n_records - 1)

    def stripped(self) -> "FastaLayout":
        """The layout of the stream the host loader produces from this file: '>' + bases per record."""
        from dataclasses import replace
        return replace(self, line_width=-1, header_fmt=">")

    def body_len(self, nb: int) -> int:
        if self.line_width < 0:
            return nb
        if self.line_width == 0:
            return nb + 1
        return nb + (nb + self.line_width - 1) // self.line_width

    def record_offsets(self):
        """(byte offset of each record [n_records+1], first global base index of each record [n_records+1])."""
        offs = np.zeros(self.n_records + 1, dtype=np.uint64)
        base0 = np.zeros(self.n_records + 1, dtype=np.uint64)
        o = b = 0
        for r in range(self.n_records):
            offs[r], base0[r] = o, b
            nb = self.record_bases(r)
            o += self.header_len + self.body_len(nb)
            b += nb
        offs[-1], base0[-1] = o, b
        return offs, base0

    @property
    def total_bytes(self) -> int:
        return int(self.record_offsets()[0][-1])


def base_letters(layout: FastaLayout, g: np.ndarray) -> np.ndarray:
    """ASCII letter of global base index g (uint64 array) under the layout's seed and masks."""
    g = g.astype(np.uint64)
    with np.errstate(over="ignore"):
        code = (splitmix64(np.uint64(layout.seed & MASK64) + g) >> np.uint64(62)).astype(np.uint8)
        out = np.frombuffer(b"ACGT", dtype=np.uint8)[code]
        if layout.soft_mask:
            slot = g // np.uint64(LC_SLOT)
            h = splitmix64((np.uint64(layout.seed & MASK64) ^ np.uint64(LC_SALT)) + slot)
            ln = (RUN_QUANTILES[(h & np.uint64(15)).astype(np.int64)] * np.uint64(3)) // np.uint64(10)
            start = (h >> np.uint64(8)) % (np.uint64(LC_SLOT) - ln)
            inside = (g % np.uint64(LC_SLOT) - start) < ln   # unsigned wrap makes "before start" huge
            out = np.where(inside, out | np.uint8(0x20), out)
        if layout.n_runs:
            slot = g // np.uint64(N_SLOT)
            h = splitmix64((np.uint64(layout.seed & MASK64) ^ np.uint64(N_SALT)) + slot)
            ln = RUN_QUANTILES[(h & np.uint64(15)).astype(np.int64)]
            start = (h >> np.uint64(8)) % (np.uint64(N_SLOT) - ln)
            inside = (g % np.uint64(N_SLOT) - start) < ln
            out = np.where(inside, np.uint8(ord("N")), out)
    return out.astype(np.uint8)


def render(layout: FastaLayout) -> np.ndarray:
    """Materialise the whole file as a uint8 array (host side; use the CUDA twin for multi-GB files)."""
    offs, base0 = layout.record_offsets()
    out = np.empty(int(offs[-1]), dtype=np.uint8)
    H, W = layout.header_len, layout.line_width
    for r in range(layout.n_records):
        o = int(offs[r])
        hdr = layout.header(r)
        assert len(hdr) == H and hdr.startswith(b">") and (W < 0 or hdr.endswith(b"\n"))
        out[o:o + H] = np.frombuffer(hdr, dtype=np.uint8)
        nb = layout.record_bases(r)
        body = out[o + H:o + H + layout.body_len(nb)]
        letters = base_letters(layout, np.arange(nb, dtype=np.uint64) + base0[r])
        if W < 0:
            body[:] = letters
        elif W == 0:
            body[:nb] = letters
            body[nb] = 10
        else:
            body[:] = 10
            idx = np.arange(nb, dtype=np.int64)
            body[idx + idx // W] = letters
    return out


def stripped_stream(layout: FastaLayout) -> np.ndarray:
    """The stream the host loader must produce from render(layout): one '>' per header, no newlines."""
    return render(layout.stripped())


# ------------------------------------------------------------------------------------------------
# the BASELINE.json configurations
# ------------------------------------------------------------------------------------------------
def config2(n_bases: int = 12_000_000, seed: int = 20142) -> FastaLayout:
    """12 Mbp yeast-scale genome, 16 records, 60-column lines (the launcher's 'full' shape)."""
    return FastaLayout(n_bases, 16, 60, ">chr%02d synthetic yeast-scale findkmer_b200\n", seed)


def config3(n_records: int = 6000, seed: int = 20143) -> FastaLayout:
    """~6000 upstream records: '>ENST%011d' + 1001 bases on ONE line (get_upstreams.pl:86-90)."""
    return FastaLayout(n_records * 1001, n_records, 0, ">ENST%011d\n", seed)


def config4(n_bases: int = 3_100_000_000, seed: int = 20144) -> FastaLayout:
    """3.1 Gbp human-scale genome, 24 records (each < 2^31 bases: the reference's seqSize is an int)."""
    return FastaLayout(n_bases, 24, 60, ">chr%02d synthetic human-scale findkmer_b200\n", seed)


def config5(n_bases: int = 3_100_000_000, seed: int = 20145) -> FastaLayout:
    """As config 4 with ~5% of positions in N runs and ~10% soft-masked lower case (both reset)."""
    return FastaLayout(n_bases, 24, 60, ">chr%02d synthetic human-scale N-runs soft-masked\n", seed,
                       n_runs=True, soft_mask=True)

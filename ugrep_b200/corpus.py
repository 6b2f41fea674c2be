"""Seeded synthetic corpora for the five BASELINE.json configs.

Every generator returns a ``numpy.uint8`` array that is LF-terminated, valid
UTF-8, has no NUL and no BOM (a BOM would make the reference transcode,
/root/reference/lib/input.cpp:664-744).  Generation is vectorised: a corpus is a
stream of vocabulary tokens, assembled with one gather.

``block(config, nbytes, seed)`` makes a line-aligned block of about ``nbytes``;
benchmarks tile a block on the device to reach the GiB sizes the configs name
(the tiled corpus is still line-aligned text, and every count scales with the
number of tiles, which the tests use as a size-independent check).
"""
from __future__ import annotations

import numpy as np

__all__ = ["block", "words_list", "CONFIG_NAMES", "english", "needles", "ing", "greek", "logs"]

CONFIG_NAMES = ("c1", "c2", "c3", "c4", "c5")

_ENGLISH = (
    "the of and to in is that it was for on are as with his they at be this from have or by one had "
    "not but what all were when we there can an your which their said if do will each about how up out "
    "them then she many some so these would other into has more her two like him see time could no make "
    "than first been its who now people my made over did down only way find use may water long little "
    "very after words called just where most know get through back much before go good new write our "
    "used me man too any day same right look think also around another came come work three word must "
    "because does part even place well such here take why things help put years different away again "
    "off went old number great tell men say small every found still between name should home big give "
    "air line set own under read last never us left end along while might next sound below saw something "
    "thought both few those always looked show large often together asked house world going want school "
    "important until form food keep children feet land side without boy once animals life enough took"
).split()


def _assemble(vocab: list[bytes], tok: np.ndarray) -> np.ndarray:
    """Concatenate vocab[tok[0]], vocab[tok[1]], ... into one uint8 array."""
    lens = np.fromiter((len(v) for v in vocab), dtype=np.int64, count=len(vocab))
    starts = np.zeros(len(vocab), dtype=np.int64)
    np.cumsum(lens[:-1], out=starts[1:])
    flat = np.frombuffer(b"".join(vocab), dtype=np.uint8)
    tl = lens[tok]
    total = int(tl.sum())
    out_start = np.zeros(len(tok), dtype=np.int64)
    np.cumsum(tl[:-1], out=out_start[1:])
    src = np.repeat(starts[tok] - out_start, tl)
    src += np.arange(total, dtype=np.int64)
    return flat[src]


def _line_layout(rng: np.random.Generator, n_lines: int, lo: int, hi: int):
    """Words per line in [lo, hi]; returns (words_per_line, is_first, is_last) over the token stream."""
    wpl = rng.integers(lo, hi + 1, size=n_lines)
    total = int(wpl.sum())
    last_idx = np.cumsum(wpl) - 1
    is_last = np.zeros(total, dtype=bool)
    is_last[last_idx] = True
    is_first = np.zeros(total, dtype=bool)
    is_first[0] = True
    is_first[last_idx[:-1] + 1] = True
    return wpl, is_first, is_last


def _variants(words: list[bytes], cap_first: bool, end: bytes):
    """vocab laid out as [mid..., first..., last...]: 'w ', 'W ', 'w<end>\\n'."""
    mid = [w + b" " for w in words]
    first = [(w[:1].upper() + w[1:] if cap_first else w) + b" " for w in words]
    last = [w + end + b"\n" for w in words]
    return mid + first + last


def _pick(n_words: int, widx: np.ndarray, is_first: np.ndarray, is_last: np.ndarray) -> np.ndarray:
    tok = widx.copy()
    tok[is_first] += n_words
    # a one-word line is both first and last: treat as last
    tok[is_last] = widx[is_last] + 2 * n_words
    return tok


def english(nbytes: int, seed: int = 42, needle: bytes | None = b"Sherlock Holmes", p_line: float = 0.002,
            extra: list[bytes] | None = None, p_extra: float = 0.0) -> np.ndarray:
    """c1 / c3 text: lines of 5-16 vocabulary words, first word capitalised, '.' terminated.

    ``needle`` is injected as a mid-line token with probability ``p_line`` per line;
    ``extra`` words (e.g. the -ing words of c3) replace a word with probability ``p_extra``.
    """
    rng = np.random.default_rng(seed)
    words = [w.encode() for w in _ENGLISH[:200]]
    if extra:
        words = words + list(extra)
    if needle is not None:
        words = words + [needle]
    nw = len(words)
    vocab = _variants(words, True, b".")
    avg_line = 10.5 * 5.3
    n_lines = max(1, int(nbytes / avg_line))
    wpl, is_first, is_last = _line_layout(rng, n_lines, 5, 16)
    total = int(wpl.sum())
    widx = rng.integers(0, 200, size=total)
    if extra and p_extra > 0:
        m = rng.random(total) < p_extra
        widx[m] = 200 + rng.integers(0, len(extra), size=int(m.sum()))
    if needle is not None and p_line > 0:
        hit = np.flatnonzero(rng.random(n_lines) < p_line)
        line_start = np.cumsum(wpl) - wpl
        # second word of the line (mid position): lines have >= 5 words
        widx[line_start[hit] + 1] = nw - 1
    return _assemble(vocab, _pick(nw, widx, is_first, is_last))


_SYLL = ("ba be bi bo bu da de di do du fa fe fi fo fu ga ge gi go gu ka ke ki ko ku la le li lo lu "
         "ma me mi mo mu na ne ni no nu ra re ri ro ru sa se si so su ta te ti to tu").split()


def _syllable_words(rng: np.random.Generator, count: int, lo: int, hi: int, taken: set[bytes],
                    avoid: set[bytes] | None = None) -> list[bytes]:
    """``count`` distinct words of lo..hi syllables that are not in ``taken``; with ``avoid``, words that contain one of
    those (syllable strings) anywhere are skipped as well."""
    out: list[bytes] = []
    while len(out) < count:
        n = int(rng.integers(lo, hi + 1))
        w = "".join(_SYLL[int(i)] for i in rng.integers(0, len(_SYLL), size=n)).encode()
        if w in taken:
            continue
        if avoid is not None and any(w[2 * i:2 * j] in avoid for i in range(n) for j in range(i + 1, n + 1)):
            continue
        taken.add(w)
        out.append(w)
    return out


def words_list(seed: int = 11, count: int = 1000) -> list[bytes]:
    """c2 words.txt: ``count`` distinct lowercase words of 2-5 syllables from a 50-syllable table, sorted
    (SURVEY.md 8(d)-2)."""
    rng = np.random.default_rng(seed)
    return sorted(_syllable_words(rng, count, 2, 5, set()))


def needles(nbytes: int, seed: int = 11, p_needle: float = 0.004, count: int = 1000, dense: bool = False) -> np.ndarray:
    """c2 text: lines of 2-12 words.

    ``dense`` (the headline corpus): words drawn uniformly from the ``count`` needles and 20 000
    distractors made the same way as the needles (2-5 syllables, no marker), so a needle may also occur inside a
    longer distractor and most word starts look like needle starts to the prefilter.
    sparse (``c2s``): a word is a needle with probability ``p_needle``, else one of 20 000 distractors of 2-3
    syllables plus a digit, so that a distractor never contains a needle (<= 3 % matching lines)."""
    rng = np.random.default_rng(seed)
    needle_words = words_list(seed, count)
    rng2 = np.random.default_rng(seed + 1)
    taken = set(needle_words)
    if dense:
        distract = _syllable_words(rng2, 20000, 2, 5, taken)
        p_needle = count / float(count + len(distract))
    else:
        base = _syllable_words(rng2, 20000, 2, 3, taken, avoid=set(needle_words))
        distract = [w + str(int(d)).encode() for w, d in zip(base, rng2.integers(0, 10, size=len(base)))]
    words = distract + needle_words
    nw = len(words)
    vocab = _variants(words, False, b"")
    avg_line = 7.0 * 7.0
    n_lines = max(1, int(nbytes / avg_line))
    wpl, is_first, is_last = _line_layout(rng, n_lines, 2, 12)
    total = int(wpl.sum())
    widx = rng.integers(0, len(distract), size=total)
    m = rng.random(total) < p_needle
    widx[m] = len(distract) + rng.integers(0, len(needle_words), size=int(m.sum()))
    return _assemble(vocab, _pick(nw, widx, is_first, is_last))


_ING = [b"Running", b"Walking", b"Thinking", b"Having", b"Looking", b"morning", b"Reading", b"Writing",
        b"evening", b"Something", b"Playing", b"king", b"Bring", b"Singing"]


def ing(nbytes: int, seed: int = 42) -> np.ndarray:
    """c3 text: the c1 generator with ~5 % -ing words."""
    return english(nbytes, seed, needle=None, extra=_ING, p_extra=0.05)


def greek(nbytes: int, seed: int = 7) -> np.ndarray:
    """c4 text: per word 10 % Greek runs (1-8 letters), 5 % naive-family, 5 % other scripts, rest ASCII."""
    rng = np.random.default_rng(seed)
    letters = [chr(c) for c in range(0x3B1, 0x3CA) if c != 0x3C2] + [chr(c) for c in range(0x391, 0x399)]
    gw = []
    for _ in range(400):
        n = int(rng.integers(1, 9))
        gw.append("".join(letters[int(i)] for i in rng.integers(0, len(letters), size=n)).encode("utf-8"))
    naive = [s.encode("utf-8") for s in ("naïve", "NAÏVE", "Naïveté", "naive", "naïvely", "NAïVE")]
    other = [s.encode("utf-8") for s in ("日本語", "Привет", "café", "über", "中文", "Здравствуйте", "señor", "façade")]
    ascii_words = [w.encode() for w in _ENGLISH[:200]]
    words = ascii_words + gw + naive + other
    nw = len(words)
    vocab = _variants(words, False, b".")
    avg_line = 9.0 * 6.0
    n_lines = max(1, int(nbytes / avg_line))
    wpl, is_first, is_last = _line_layout(rng, n_lines, 4, 14)
    total = int(wpl.sum())
    r = rng.random(total)
    widx = rng.integers(0, len(ascii_words), size=total)
    g = r < 0.10
    widx[g] = len(ascii_words) + rng.integers(0, len(gw), size=int(g.sum()))
    v = (r >= 0.10) & (r < 0.15)
    widx[v] = len(ascii_words) + len(gw) + rng.integers(0, len(naive), size=int(v.sum()))
    o = (r >= 0.15) & (r < 0.20)
    widx[o] = len(ascii_words) + len(gw) + len(naive) + rng.integers(0, len(other), size=int(o.sum()))
    return _assemble(vocab, _pick(nw, widx, is_first, is_last))


def logs(nbytes: int, seed: int = 7) -> np.ndarray:
    """c5 text: ``ISO-timestamp LEVEL svcN call 555-dddd ext ddd-dddd id=n`` log lines."""
    rng = np.random.default_rng(seed)
    line_len = 72.0
    n = max(1, int(nbytes / line_len))
    levels = [b"INFO", b"DEBUG", b"ERROR", b"WARN", b"TRACE"]
    d2 = [b"%02d" % i for i in range(100)]
    d3 = [b"%03d" % i for i in range(1000)]
    d4 = [b"%04d" % i for i in range(10000)]
    vocab: list[bytes] = []
    off = {}
    for name, lst in (("d2", d2), ("d3", d3), ("d4", d4), ("lvl", levels)):
        off[name] = len(vocab)
        vocab += lst
    fixed = [b"2026-", b"-", b"T", b":", b" ", b" svc", b" call 555-", b" ext ", b" id=", b"\n"]
    off["fx"] = len(vocab)
    vocab += fixed
    fx = off["fx"]
    cols = [
        np.full(n, fx + 0), off["d2"] + rng.integers(1, 13, size=n), np.full(n, fx + 1),
        off["d2"] + rng.integers(1, 29, size=n), np.full(n, fx + 2), off["d2"] + rng.integers(0, 24, size=n),
        np.full(n, fx + 3), off["d2"] + rng.integers(0, 60, size=n), np.full(n, fx + 3),
        off["d2"] + rng.integers(0, 60, size=n), np.full(n, fx + 4), off["lvl"] + rng.integers(0, 5, size=n),
        np.full(n, fx + 5), off["d2"] + rng.integers(0, 100, size=n), np.full(n, fx + 6),
        off["d4"] + rng.integers(0, 10000, size=n), np.full(n, fx + 7), off["d3"] + rng.integers(0, 1000, size=n),
        np.full(n, fx + 1), off["d4"] + rng.integers(0, 10000, size=n), np.full(n, fx + 8),
        off["d4"] + rng.integers(0, 10000, size=n), np.full(n, fx + 9),
    ]
    tok = np.stack(cols, axis=1).reshape(-1)
    return _assemble(vocab, tok)


def block(config: str, nbytes: int, seed: int | None = None) -> np.ndarray:
    """A line-aligned block of about ``nbytes`` for config c1..c5 (c2 = the dense variant, c2s = the sparse one)."""
    if config == "c1":
        return english(nbytes, 42 if seed is None else seed)
    if config == "c2":
        return needles(nbytes, 11 if seed is None else seed, dense=True)
    if config == "c2s":
        return needles(nbytes, 11 if seed is None else seed)
    if config == "c3":
        return ing(nbytes, 42 if seed is None else seed)
    if config == "c4":
        return greek(nbytes, 7 if seed is None else seed)
    if config == "c5":
        return logs(nbytes, 7 if seed is None else seed)
    raise ValueError("unknown config %r" % (config,))

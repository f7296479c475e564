"""The `standard` analyzer of the reference's text fields: UAX#29 word segmentation + lower-casing, no stop words.

The reference maps `unstructuredText`, `patientName`, `observationValue` ... as {"type": "text"} with no analyzer
(app/main.py:361-561), so OpenSearch 2.11 / Lucene 9.7 applies StandardTokenizer (the Unicode Text Segmentation word
break rules, Unicode 9.0 tables) followed by LowerCaseFilter.  What that means for clinical text:

  * letters and digits run together ("hba1c"), WB5 / WB8-10;
  * a single MidLetter / MidNumLet / apostrophe BETWEEN two letters stays inside the token ("o'neil", "e.g", "x:y",
    "example.com"), WB6-7; a single MidNum / MidNumLet / apostrophe BETWEEN two digits does too ("3.5", "1,000"),
    WB11-12 -- but "/" and "-" are neither, so "120/80" and "2021-03-04" split;
  * "_" (ExtendNumLet) joins ("sars_cov_2"), WB13a-b;
  * combining marks, ZWJ and format characters attach to what precedes them (WB4), so a decomposed "e" + U+0301 keeps its accent;
  * Han ideographs and Hiragana come out one character per token, a run of Katakana or of a South-East-Asian script
    (Thai, Lao, Myanmar, Khmer: no spaces between words) as one token;
  * tokens longer than 255 characters are cut into 255-character pieces (StandardAnalyzer's maxTokenLength);
  * lower-casing is per code point (Character.toLowerCase): no final-sigma context, U+0130 -> "i".

Not restated: emoji tokens (Lucene 9 emits them; no text field of the reference's mapping is expected to hold any).
ASCII text -- almost everything the reference ingests -- takes a compiled regular expression that states the same
rules for the ASCII classes; anything else goes through the explicit scanner below.
"""
from __future__ import annotations

import re
import unicodedata

MAX_TOKEN_LENGTH = 255

# word-break classes
OTHER, ALETTER, HEBREW, NUMERIC, KATAKANA, EXTNUMLET, MIDLETTER, MIDNUMLET, MIDNUM, SQUOTE, DQUOTE, EXTEND, \
    IDEOGRAPHIC, HIRAGANA, COMPLEX = range(15)

_MIDLETTER = frozenset("\u003a\u00b7\u0387\u05f4\u2027\ufe13\ufe55\uff1a")
_MIDNUMLET = frozenset("\u002e\u2018\u2019\u2024\ufe52\uff07\uff0e")
_MIDNUM = frozenset("\u002c\u003b\u037e\u0589\u060c\u060d\u066c\u07f8\u2044\ufe10\ufe14\ufe50\ufe54\uff0c\uff1b")
_NOT_FORMAT = frozenset("\u200b")          # ZWSP is Cf but not WB=Format


def _in(cp: int, ranges) -> bool:
    return any(lo <= cp <= hi for lo, hi in ranges)


_IDEO = ((0x4E00, 0x9FFF), (0x3400, 0x4DBF), (0x20000, 0x2A6DF), (0x2A700, 0x2EBEF), (0xF900, 0xFAFF),
         (0x2F800, 0x2FA1F), (0x3007, 0x3007), (0x3021, 0x3029), (0x3038, 0x303A))
_HIRA = ((0x3041, 0x3096), (0x309D, 0x309F))
_KATA = ((0x30A0, 0x30FA), (0x30FC, 0x30FF), (0x31F0, 0x31FF), (0x32D0, 0x32FE), (0x3300, 0x3357), (0xFF66, 0xFF9D),
         (0x3031, 0x3035), (0x309B, 0x309C), (0x1B000, 0x1B000))
_HEBR = ((0x05D0, 0x05EA), (0x05F0, 0x05F2), (0xFB1D, 0xFB1D), (0xFB1F, 0xFB28), (0xFB2A, 0xFB4F))
# Line_Break = Complex_Context (SA): scripts written without spaces between words
_SA = ((0x0E00, 0x0E7F), (0x0E80, 0x0EFF), (0x1000, 0x109F), (0x1780, 0x17FF), (0x1950, 0x197F), (0x1980, 0x19DF),
       (0x1A20, 0x1AAF), (0xA9E0, 0xA9FF), (0xAA60, 0xAADF))

_cache: dict[str, int] = {}


def word_break_class(ch: str) -> int:
    c = _cache.get(ch)
    if c is not None:
        return c
    cp = ord(ch)
    cat = unicodedata.category(ch)
    if ch == "'":
        c = SQUOTE
    elif ch == '"':
        c = DQUOTE
    elif ch in _MIDLETTER:
        c = MIDLETTER
    elif ch in _MIDNUMLET:
        c = MIDNUMLET
    elif ch in _MIDNUM:
        c = MIDNUM
    elif cat == "Nd" or ch == "\u066b":
        c = NUMERIC
    elif _in(cp, _SA) and cat[0] in "LM":
        c = COMPLEX                          # before Extend: the vowel signs of these scripts stay in the run
    elif cat in ("Mn", "Me", "Mc") or ch in "\u200c\u200d" or (cat == "Cf" and ch not in _NOT_FORMAT):
        c = EXTEND                           # Extend | Format | ZWJ: all absorbed by WB4
    elif cat == "Pc":
        c = EXTNUMLET
    elif _in(cp, _IDEO):
        c = IDEOGRAPHIC
    elif _in(cp, _HIRA):
        c = HIRAGANA
    elif _in(cp, _KATA):
        c = KATAKANA
    elif _in(cp, _HEBR):
        c = HEBREW
    elif cat[0] == "L" or cat == "Nl" or 0x24B6 <= cp <= 0x24E9:
        c = ALETTER
    else:
        c = OTHER
    if len(_cache) < 65536:
        _cache[ch] = c
    return c


def _lower(tok: str) -> str:
    if tok.isascii():
        return tok.lower()
    out = []
    for ch in tok:
        lo = ch.lower()
        out.append(lo if len(lo) == 1 else lo[0])        # Character.toLowerCase: one code point in, one out
    return "".join(out)


def _emit(out: list, tok: str):
    for i in range(0, len(tok), MAX_TOKEN_LENGTH):
        out.append(_lower(tok[i:i + MAX_TOKEN_LENGTH]))


# the same rules for ASCII input: units are letters, digits and "_"; one of : . ' between two letters, one of , ; . '
# between two digits; a token needs a letter or a digit
_ASCII_TOKEN = re.compile(r"(?:[A-Za-z](?:[:.'](?=[A-Za-z]))?|[0-9](?:[.,;'](?=[0-9]))?|_)+")
_ASCII_CORE = re.compile(r"[A-Za-z0-9]")


def _analyze_ascii(text: str) -> list[str]:
    # lower-casing first does not move a boundary (a letter stays a letter), and it is one C call instead of one per token
    toks = _ASCII_TOKEN.findall(text.lower())
    if "_" in text:                                   # a run of underscores alone is not a token
        toks = [t for t in toks if t[0] != "_" or _ASCII_CORE.search(t)]
    if len(text) > MAX_TOKEN_LENGTH and any(len(t) > MAX_TOKEN_LENGTH for t in toks):
        out: list[str] = []
        for t in toks:
            if len(t) <= MAX_TOKEN_LENGTH:
                out.append(t)
            else:
                _emit(out, t)
        return out
    return toks


_LETTER = (ALETTER, HEBREW)


def _joins(prev: int, cur: int) -> bool:
    if cur == EXTNUMLET:                                       # WB13a
        return prev in (ALETTER, HEBREW, NUMERIC, KATAKANA, EXTNUMLET)
    if prev == EXTNUMLET:                                      # WB13b
        return cur in (ALETTER, HEBREW, NUMERIC, KATAKANA)
    if prev in _LETTER:
        return cur in _LETTER or cur == NUMERIC                # WB5, WB9
    if prev == NUMERIC:
        return cur == NUMERIC or cur in _LETTER                # WB8, WB10
    if prev == KATAKANA:
        return cur == KATAKANA                                 # WB13
    return False


def _analyze_unicode(text: str) -> list[str]:
    n = len(text)
    cls = [word_break_class(ch) for ch in text]
    out: list[str] = []

    def skip_ext(j: int) -> int:                               # WB4: X (Extend | Format | ZWJ)* -> X
        while j < n and cls[j] == EXTEND:
            j += 1
        return j

    i = 0
    while i < n:
        c = cls[i]
        if c in (IDEOGRAPHIC, HIRAGANA):
            j = skip_ext(i + 1)
            _emit(out, text[i:j])
            i = j
        elif c == COMPLEX:
            j = i + 1
            while j < n and cls[j] in (COMPLEX, EXTEND):
                j += 1
            _emit(out, text[i:j])
            i = j
        elif c in (ALETTER, HEBREW, NUMERIC, KATAKANA, EXTNUMLET):
            start, prev, core = i, c, c != EXTNUMLET
            j = skip_ext(i + 1)
            while j < n:
                d = cls[j]
                if _joins(prev, d):
                    prev = d
                    core = core or d != EXTNUMLET
                    j = skip_ext(j + 1)
                    continue
                if d in (MIDLETTER, MIDNUMLET, SQUOTE, MIDNUM, DQUOTE):
                    k = skip_ext(j + 1)
                    nxt = cls[k] if k < n else OTHER
                    if prev in _LETTER and d in (MIDLETTER, MIDNUMLET, SQUOTE) and nxt in _LETTER:    # WB6-7
                        prev, j = nxt, skip_ext(k + 1)
                        continue
                    if prev == NUMERIC and d in (MIDNUM, MIDNUMLET, SQUOTE) and nxt == NUMERIC:       # WB11-12
                        prev, j = nxt, skip_ext(k + 1)
                        continue
                    if prev == HEBREW and d == DQUOTE and nxt == HEBREW:                              # WB7b-c
                        prev, j = nxt, skip_ext(k + 1)
                        continue
                    if prev == HEBREW and d == SQUOTE:                                                # WB7a
                        j = k
                        prev = OTHER
                        continue
                break
            if core:
                _emit(out, text[start:j])
            i = j
        else:
            i += 1
    return out


def analyze(text: str) -> list[str]:
    """Tokens of `text` as the reference's `standard` analyzer produces them (see the module docstring)."""
    if not text:
        return []
    return _analyze_ascii(text) if text.isascii() else _analyze_unicode(text)

"""Analyzer restatement (test infrastructure).

The reference maps its text fields as {"type": "text"} with no analyzer (app/main.py:361-561), i.e. OpenSearch's
`standard` analyzer: Lucene StandardTokenizer (UAX#29 word break rules, Unicode 9.0 property tables) + LowerCaseFilter,
no stop words, maxTokenLength 255 (third-party, UNPINNED: restated from the published annex; the known answers in
tests/golden/analyzer_cases.json were worked out by hand from the rules, WB numbers beside each).

Written independently of rassengine_b200/analysis.py (which scans characters with an explicit state machine): here every
character is mapped to a one-letter class, Extend/Format/ZWJ characters are folded into the unit they follow (WB4), and
ONE regular expression over the class string states which units stay together:

    A ALetter   H Hebrew_Letter   N Numeric   K Katakana   U ExtendNumLet   L MidLetter   B MidNumLet   M MidNum
    Q Single_Quote   D Double_Quote   I ideograph   R Hiragana   S South-East-Asian (complex context)   O anything else

    word      = ( [AH] ( [LBQ] before [AH] | D between two H )?  |  N ( [MBQ] before N )? )+          WB5-12
    group     = word | K+                                                                               WB13
    token     = U* group ( U+ group )* ( H-final Q | U* )                                               WB13a/b, WB7a
              | I | R | S+                                   (ideographs / Hiragana: one per token; SA runs: whole)

The synthetic corpora only emit lowercase [a-z0-9]+ tokens, so `str.split()` (the reference's own chunker,
app/main.py:2160-2170) agrees with this on them.
"""
from __future__ import annotations

import re
import unicodedata as ud

MAX_TOKEN_LENGTH = 255

_MID = {}
for _cp in (0x3A, 0xB7, 0x387, 0x5F4, 0x2027, 0xFE13, 0xFE55, 0xFF1A):
    _MID[_cp] = "L"
for _cp in (0x2E, 0x2018, 0x2019, 0x2024, 0xFE52, 0xFF07, 0xFF0E):
    _MID[_cp] = "B"
for _cp in (0x2C, 0x3B, 0x37E, 0x589, 0x60C, 0x60D, 0x66C, 0x7F8, 0x2044, 0xFE10, 0xFE14, 0xFE50, 0xFE54, 0xFF0C, 0xFF1B):
    _MID[_cp] = "M"
_MID[0x27] = "Q"
_MID[0x22] = "D"

_BLOCKS = [  # (first, last, class) for the scripts that do not follow "letter = ALetter"
    (0x0E00, 0x0EFF, "S"), (0x1000, 0x109F, "S"), (0x1780, 0x17FF, "S"), (0x1950, 0x19DF, "S"), (0x1A20, 0x1AAF, "S"),
    (0xA9E0, 0xA9FF, "S"), (0xAA60, 0xAADF, "S"),
    (0x3041, 0x3096, "R"), (0x309D, 0x309F, "R"),
    (0x3031, 0x3035, "K"), (0x309B, 0x309C, "K"), (0x30A0, 0x30FA, "K"), (0x30FC, 0x30FF, "K"), (0x31F0, 0x31FF, "K"),
    (0x32D0, 0x32FE, "K"), (0x3300, 0x3357, "K"), (0xFF66, 0xFF9D, "K"), (0x1B000, 0x1B000, "K"),
    (0x3007, 0x3007, "I"), (0x3021, 0x3029, "I"), (0x3038, 0x303A, "I"), (0x3400, 0x4DBF, "I"), (0x4E00, 0x9FFF, "I"),
    (0xF900, 0xFAFF, "I"), (0x20000, 0x2A6DF, "I"), (0x2A700, 0x2EBEF, "I"), (0x2F800, 0x2FA1F, "I"),
    (0x05D0, 0x05EA, "H"), (0x05F0, 0x05F2, "H"), (0xFB1D, 0xFB1D, "H"), (0xFB1F, 0xFB28, "H"), (0xFB2A, 0xFB4F, "H"),
]


def char_class(ch: str) -> str:
    cp = ord(ch)
    if cp in _MID:
        return _MID[cp]
    gc = ud.category(ch)
    if gc == "Nd" or cp == 0x66B:
        return "N"
    block = next((c for lo, hi, c in _BLOCKS if lo <= cp <= hi), None)
    if block == "S":
        if gc[0] in "LM":
            return "S"
        block = None
    if gc[0] == "M" or cp in (0x200C, 0x200D) or (gc == "Cf" and cp != 0x200B):
        return "E"
    if gc == "Pc":
        return "U"
    if block:
        return block
    if gc[0] == "L" or gc == "Nl" or 0x24B6 <= cp <= 0x24E9:
        return "A"
    return "O"


_WORD = r"(?:[AH](?:[LBQ](?=[AH])|D(?<=HD)(?=H))?|N(?:[MBQ](?=N))?)+"
_GROUP = rf"(?:{_WORD}|K+)"
_TOKEN = re.compile(rf"U*{_GROUP}(?:U+{_GROUP})*(?:(?<=H)Q(?![AH])|U*)|I|R|S+")


def _lower(tok: str) -> str:
    """LowerCaseFilter: Character.toLowerCase per code point (no context rules, one code point out)."""
    return "".join((c.lower() if len(c.lower()) == 1 else c.lower()[0]) for c in tok)


def analyze(text: str) -> list[str]:
    if not text:
        return []
    # WB4: fold Extend / Format / ZWJ into the unit they follow (one at the very start stands alone as "other")
    classes: list[str] = []
    starts: list[int] = []
    for i, ch in enumerate(text):
        c = char_class(ch)
        if c == "E" and classes:
            continue
        classes.append("O" if c == "E" else c)
        starts.append(i)
    starts.append(len(text))
    out: list[str] = []
    for m in _TOKEN.finditer("".join(classes)):
        tok = text[starts[m.start()]:starts[m.end()]]
        for o in range(0, len(tok), MAX_TOKEN_LENGTH):
            out.append(_lower(tok[o:o + MAX_TOKEN_LENGTH]))
    return out

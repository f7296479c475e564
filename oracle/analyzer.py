"""Analyzer restatement (test infrastructure).

The reference maps `unstructuredText` as {"type": "text"} with no analyzer
(app/main.py:555-556), i.e. OpenSearch's `standard` analyzer: UAX#29 word
segmentation + lowercase, no stop words (third-party, UNPINNED).  For ASCII
input that is lower-casing and splitting on runs of non-alphanumerics, which is
what is restated here; the synthetic corpora only emit lowercase [a-z0-9]+ tokens
so `str.split()` (the reference's own chunker, app/main.py:2160-2170) agrees.
"""
from __future__ import annotations

import re

_TOKEN = re.compile(r"[0-9a-z]+")


def analyze(text: str) -> list[str]:
    return _TOKEN.findall(text.lower())

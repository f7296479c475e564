"""`opensearchpy`-compatible import shim: put `rassengine_b200/shim` on PYTHONPATH and the reference's
`from opensearchpy import OpenSearch, RequestsHttpConnection` / `from opensearchpy.helpers import bulk`
(app/main.py:31-32) resolve to the in-process B200 engine."""
from rassengine_b200.client import B200Client as OpenSearch, NotFoundError, RequestError  # noqa: F401
from . import helpers  # noqa: F401


class RequestsHttpConnection:      # accepted as connection_class=..., unused
    pass


__all__ = ["OpenSearch", "RequestsHttpConnection", "NotFoundError", "RequestError", "helpers"]

from rassengine_b200.client import bulk  # noqa: F401

"""Bindings of the device executor INTO the reference package (``import tensorsat`` must work)."""

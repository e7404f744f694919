"""Declarative derived points (compiled to device ops by core/topology.py)."""

"""Configuration records parsed from the geometry mapping."""

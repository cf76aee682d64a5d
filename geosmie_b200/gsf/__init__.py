"""Generalized-spherical-function moments of an optics table (mirror of src/gsf/)."""

"""Drop-in replacements of the reference's `pnn` package entry points (inference only)."""

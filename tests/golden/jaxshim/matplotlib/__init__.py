"""matplotlib stand-in (the reference imports pyplot at module scope for its __main__ plots only)."""

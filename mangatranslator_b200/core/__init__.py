"""Reference-facing stage interfaces (mirror of the reference's `core` package for the vision hot path only)."""

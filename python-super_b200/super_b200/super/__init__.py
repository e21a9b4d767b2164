"""Drop-in counterparts of the reference's `super/` package (same module and class names)."""

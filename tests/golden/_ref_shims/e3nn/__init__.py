"""Import stub for e3nn (used only by a descriptor family that is not evaluated here)."""

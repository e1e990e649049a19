"""Import stub for wcmatch (deepmd.utils.path imports wcmatch.glob at module level; no globbing happens here)."""

"""Test infrastructure (CPU oracle + reference bindings). Never imported by panmap_b200/."""

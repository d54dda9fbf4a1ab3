"""Host-side mirror of the reference package `vispec.model` (same module and symbol names)."""

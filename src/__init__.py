"""Mirror of the reference's ``src`` package root (all of its sub-packages are 0-byte files upstream)."""

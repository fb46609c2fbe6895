"""pyvista stand-in (import only).  Test infrastructure only."""

"""Host-side mirror of the reference's ``yolo`` package (config + network builder)."""

"""Import stand-in (yolo_head/exportable_mesh_model.py imports it at module level; export() is never run here).
TEST INFRASTRUCTURE ONLY."""

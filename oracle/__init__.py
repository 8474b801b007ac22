"""CPU oracle for the vello compute pipeline -- TEST INFRASTRUCTURE ONLY (see oracle/vbo.h)."""

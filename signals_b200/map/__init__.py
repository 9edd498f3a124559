"""Placeholder for ``signals.map`` (grid map, command language, undo/redo: out of scope, SURVEY.md 2 rows
11-12).  Only the import path exists so scripts that ``import signals.map.control`` keep importing; patch
files are replayed by ``signals_b200.sigs``."""

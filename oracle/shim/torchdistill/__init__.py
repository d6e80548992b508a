"""oracle/shim/torchdistill -- import-level stand-ins for the torchdistill symbols that
`sc2bench.models` / `sc2bench.analysis` need (SURVEY.md 8b last row).  TEST INFRASTRUCTURE ONLY:
torchdistill is not installable here; these exist so the reference's own model code can be imported
to generate golden vectors.  Behaviour restated from memory of torchdistill 1.x."""

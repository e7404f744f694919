"""Result tables of solved sweeps (wide-form files, batch Arrow sink)."""

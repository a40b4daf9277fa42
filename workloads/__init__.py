"""Synthetic windows and sequences for tests, smoke() and bench.py (SURVEY.md §8d recipe).

Test / benchmark infrastructure: nothing under photobundle_b200/ (the product package) imports this.
"""

"""nh3 absorption formalisms."""

"""clouds absorption formalisms."""

"""h2o absorption formalisms."""

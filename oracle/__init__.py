"""ORACLE — test infrastructure only (see each module's header). Not part of the product."""

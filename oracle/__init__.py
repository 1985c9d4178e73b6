"""TEST INFRASTRUCTURE ONLY — CPU oracle and compiled-reference bindings.  Never imported by gsdr_b200."""

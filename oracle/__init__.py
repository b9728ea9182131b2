"""CPU oracle: TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header). Never imported by the
product package `devito_b200`."""

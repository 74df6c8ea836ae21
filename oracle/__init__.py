"""TEST INFRASTRUCTURE: CPU parity oracle for the SPH step.  Never imported by the product package."""

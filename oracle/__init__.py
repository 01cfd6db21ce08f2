"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see psnode_oracle.py).  Never imported by the product packages."""

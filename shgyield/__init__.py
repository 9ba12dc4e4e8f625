"""Import path of the reference package (`import shgyield.shg`), served by `shgyield_b200`."""

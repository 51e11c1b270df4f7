"""CPU oracle of the reconstruction hot path -- test infrastructure, never imported by the product."""

"""loans_b200 -- B200-native STN crop stage of LoANs (placeholder, filled in below)."""

"""Mirror of the reference package of the same name (B200-native implementation)."""

"""B200-native QAOA statevector engine behind the OpenQuantumComputing/QAOA API."""

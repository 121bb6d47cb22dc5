"""Secondary workloads of bench.py (`--workload c5`, `--workload c3`): measurement harness, not part of the product
package -- their `cpu_baseline` legs are, next to bench.py's own, the only places outside tests/ that call oracle/."""

"""Metric / Statistic surface of the B200 engine (mirrors weatherbenchX.metrics)."""

from diffuscene_b200.stats_logger import AverageAggregator, StatsLogger  # noqa: F401

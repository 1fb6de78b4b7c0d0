"""Seeded link-update test networks shared by the golden generator, the oracle tests and the GPU parity tests."""
from mezzo_b200 import synth

# name -> kwargs of synth.sim_grid. "det" cases use deterministic/const servers: the GPU must match BIT-EXACTLY;
# stochastic cases go through log/sqrt/sin (CUDA libm vs glibc differ by <= 2 ulp): tolerance 1e-6 relative.
CASES = {
    "freeflow_small": dict(nx=6, ny=5, seed=1, od_every=2, n_od=12, trips_per_hour=300, horizon=1800.0),
    "congested_short": dict(nx=6, ny=5, seed=2, od_every=2, n_od=20, trips_per_hour=1500, horizon=1800.0, len_lo=60,
                            len_hi=200, connector_len=50),
    "two_lane_lookback3": dict(nx=8, ny=8, seed=3, od_every=3, n_od=40, trips_per_hour=900, horizon=1800.0,
                               two_lane_prob=0.5, lookback=3),
    "det_servers": dict(nx=10, ny=10, seed=5, od_every=3, n_od=60, trips_per_hour=400, horizon=3600.0, server_type=2,
                        server_mu=2.0, server_sd=0.0, dest_server=(2, 0.5, 0.0, 0.0)),
    "det_gridlock": dict(nx=5, ny=5, seed=4, od_every=2, n_od=16, trips_per_hour=2500, horizon=900.0, len_lo=40,
                         len_hi=90, connector_len=30, server_type=2, server_mu=1.5, server_sd=0.0,
                         dest_server=(2, 0.4, 0.0, 0.0)),
    "const_servers_dynamit": dict(nx=7, ny=6, seed=6, od_every=2, n_od=30, trips_per_hour=700, horizon=1800.0,
                                  server_type=0, server_mu=0, server_sd=0, sd_type=1, sd_alpha=1.5, sd_beta=2.0,
                                  dest_server=(0, 0, 0, 0)),
    "det_two_lane_congested": dict(nx=6, ny=6, seed=8, od_every=2, n_od=30, trips_per_hour=1800, horizon=1200.0,
                                   len_lo=50, len_hi=150, connector_len=40, two_lane_prob=0.4, lookback=4,
                                   server_type=2, server_mu=1.2, server_sd=0.0, dest_server=(2, 0.3, 0.0, 0.0)),
    "const_servers_linear": dict(nx=6, ny=6, seed=9, od_every=2, n_od=24, trips_per_hour=600, horizon=1800.0,
                                 server_type=0, server_mu=0, server_sd=0, dest_server=(0, 0, 0, 0)),
    # all-integer parameters: equal link lengths and deterministic 1 s / 2 s servers make thousands of events fall on
    # EXACTLY equal times, so the result depends on the reference's FIFO order among equal-time events
    "det_ties_small": dict(nx=6, ny=6, seed=12, od_every=2, n_od=30, trips_per_hour=1800, horizon=1200.0, server_type=2,
                           server_mu=1.0, server_sd=0.0, dest_server=(2, 1.0, 0, 0), len_lo=60, len_hi=60,
                           connector_len=30, two_lane_prob=0.5, vfree=15.0),
    "det_ties_grid": dict(nx=8, ny=8, seed=11, od_every=2, n_od=60, trips_per_hour=1200, horizon=1800.0, server_type=2,
                          server_mu=2.0, server_sd=0.0, dest_server=(2, 1.0, 0, 0), len_lo=150, len_hi=150,
                          connector_len=150, two_lane_prob=0.3),
    # route choice: up to 3 routes per OD pair, time-varying historical link times, Kirchhoff shares
    "route_choice_det": dict(nx=7, ny=7, seed=21, od_every=2, n_od=40, trips_per_hour=500, horizon=3600.0, server_type=2,
                             server_mu=1.5, server_sd=0.0, dest_server=(2, 0.5, 0.0, 0.0), alt_routes=3,
                             hist_variation=0.6, n_periods=6),
    # negative-exponential OD servers (od_servers_deterministic=0): Poisson arrivals drawn from each ODaction's own Random
    "poisson_demand": dict(nx=6, ny=6, seed=23, od_every=2, n_od=30, trips_per_hour=700, horizon=1800.0, alt_routes=2,
                           hist_variation=0.3, od_servers_deterministic=0, randseed=777),
    # server-rate changes (serverrates.dat → ChangeRateAction): 60 turning servers change mu and/or sd
    "rate_changes_det": dict(nx=7, ny=6, seed=25, od_every=2, n_od=30, trips_per_hour=900, horizon=1800.0, server_type=2,
                             server_mu=1.5, server_sd=0.0, dest_server=(2, 0.5, 0.0, 0.0), rate_changes=60),
    "rate_changes_stoch": dict(nx=6, ny=6, seed=26, od_every=2, n_od=30, trips_per_hour=900, horizon=1800.0,
                               rate_changes=60),
    # delete_bad_routes=1: ODpair::delete_spurious_routes prunes and re-orders the route sets before the run (row f2) and
    # leaves the Route::cost caches primed, which changes the route choice of the whole run
    "route_pruning_det": dict(nx=7, ny=7, seed=27, od_every=2, n_od=40, trips_per_hour=500, horizon=3600.0, server_type=2,
                              server_mu=1.5, server_sd=0.0, dest_server=(2, 0.5, 0.0, 0.0), alt_routes=4,
                              hist_variation=0.9, n_periods=6, delete_bad_routes=1, max_rel_route_cost=1.12,
                              small_od_rate=300.0),
    # Parameters::server_type = 3 / 4: every class-Server object (type-1 turning servers, the destinations' default servers)
    # draws log-normal / log-logistic headways (B/server.cpp:32-50); file type 4 = LogLogisticDelayServer
    "lognormal_servers": dict(nx=6, ny=6, seed=28, od_every=2, n_od=30, trips_per_hour=700, horizon=1800.0,
                              server_mu=1.0, server_sd=0.3, server_delay=0.8, param_server_type=3,
                              dest_server=(1, 0.3, 0.1, 0.2)),  # (a zero delay makes lnrandom(0, sd) NaN in the reference)
    "loglogistic_servers": dict(nx=6, ny=6, seed=29, od_every=2, n_od=30, trips_per_hour=700, horizon=1800.0,
                                server_mu=1.8, server_sd=6.0, param_server_type=4, dest_server=(1, 0.5, 6.0, 0.0)),
    "loglogistic_file_type": dict(nx=6, ny=5, seed=30, od_every=2, n_od=24, trips_per_hour=700, horizon=1800.0,
                                  server_type=4, server_mu=1.8, server_sd=7.0),
    # traffic signals (signal.dat): 8 signal controls with two plans each (plan switch, plan end)
    "signals_det": dict(nx=6, ny=6, seed=31, od_every=2, n_od=30, trips_per_hour=600, horizon=1800.0, server_type=2,
                        server_mu=1.5, server_sd=0.0, dest_server=(2, 0.5, 0.0, 0.0), signals=8),
    "signals_stoch": dict(nx=6, ny=6, seed=32, od_every=2, n_od=30, trips_per_hour=600, horizon=1800.0, signals=8),
    # demand slices (demand.dat "slices:", row f1): MatrixActions change the OD rates four times — pairs speed up, slow down,
    # stop for good (rate 0) and pairs that start inactive (rate 0 / 0.5) are switched on; route choice over 2 routes
    "demand_slices_det": dict(nx=6, ny=6, seed=33, od_every=2, n_od=30, trips_per_hour=500, horizon=2400.0, server_type=2,
                              server_mu=1.5, server_sd=0.0, dest_server=(2, 0.5, 0.0, 0.0), demand_slices=4, alt_routes=2,
                              hist_variation=0.3),
    "demand_slices_poisson": dict(nx=6, ny=6, seed=34, od_every=2, n_od=30, trips_per_hour=500, horizon=2400.0,
                                  demand_slices=3, od_servers_deterministic=0, randseed=4242),
    # incidents without information broadcast (incident file, row f4): slow sdfuncs replace the links' own on the four busiest
    # links for a part of the horizon, two incidents overlap on one link; the stop also resets a spill-back state
    "incidents_det": dict(nx=6, ny=6, seed=35, od_every=2, n_od=30, trips_per_hour=1100, horizon=2400.0, server_type=2,
                          server_mu=1.5, server_sd=0.0, dest_server=(2, 0.5, 0.0, 0.0), incidents=4, len_lo=80, len_hi=220,
                          connector_len=60),
    "incidents_stoch": dict(nx=6, ny=6, seed=36, od_every=2, n_od=30, trips_per_hour=900, horizon=2400.0, incidents=3,
                            two_lane_prob=0.3, lookback=4),
    "route_choice_stoch": dict(nx=6, ny=6, seed=22, od_every=2, n_od=30, trips_per_hour=800, horizon=1800.0, alt_routes=3,
                               hist_variation=0.4),
}
DET = {"incidents_det", "demand_slices_det", "signals_det", "route_pruning_det", "rate_changes_det", "route_choice_det", "det_servers", "det_gridlock", "det_two_lane_congested", "const_servers_linear", "det_ties_small", "det_ties_grid", "vehicle_mix_det"}
# several vehicle classes (vehiclemix.dat; probabilities that do not sum to 1 are normalised by Vtypes::initialize): the class —
# and with it the length that queue space, densities and shock waves see — is drawn per trip from one network-wide stream
CASES["vehicle_mix_det"] = dict(nx=6, ny=6, seed=41, od_every=2, n_od=30, trips_per_hour=1500, horizon=1500.0, len_lo=60,
                                len_hi=180, connector_len=50, two_lane_prob=0.3, lookback=5, server_type=2, server_mu=1.3,
                                server_sd=0.0, dest_server=(2, 0.4, 0.0, 0.0),
                                vehicle_classes=[(0.55, 5.0), (0.25, 9.5), (0.15, 16.0)])
CASES["vehicle_mix_stoch"] = dict(nx=6, ny=5, seed=42, od_every=2, n_od=24, trips_per_hour=1200, horizon=1500.0, len_lo=70,
                                  len_hi=220, connector_len=60, od_servers_deterministic=0,
                                  vehicle_classes=[(0.5, 4.5), (0.5, 12.0)])

GOLDEN = ["freeflow_small", "congested_short", "det_gridlock", "det_two_lane_congested", "const_servers_linear",
          "det_ties_small", "route_choice_det", "route_choice_stoch", "poisson_demand", "rate_changes_det",
          "rate_changes_stoch", "route_pruning_det", "lognormal_servers", "loglogistic_servers", "loglogistic_file_type",
          "signals_det", "signals_stoch", "demand_slices_det", "demand_slices_poisson", "incidents_det",
          "incidents_stoch", "vehicle_mix_det", "vehicle_mix_stoch"]


# event counts are not comparable between the oracle (one event per Stage / SignalPlan / SignalControl action, like the
# reference) and the engine (one event per precomputed signal event)
SIGNALS = {"signals_det", "signals_stoch"}


def make(name):
    return synth.sim_grid(**CASES[name])

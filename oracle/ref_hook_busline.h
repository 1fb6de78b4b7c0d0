// TEST INFRASTRUCTURE: force-included ONLY when compiling the reference's busline.cpp in place (after busline.h, so the
// declarations are untouched). Two call sites get a logging macro, the sources are not changed:
//   * Busstop::execute, "bus enters the stop" (B/busline.cpp:1279-1281): entering_trip->set_enter_time(time) right after
//     exit_time = calc_exiting_time(...) — logs (vehicle, trip, stop, link, time the bus reaches the stop, time it leaves)
//     with full precision (o_transitlog_out.dat prints six digits);
//   * Bustrip::activate (B/busline.cpp:798): origin->insert_veh(busv, dispatch_time) — logs the time the bus enters the network and the
//     length of the vehicle that runs the trip AT THAT MOMENT (bus vehicles are shared between trips and reset when a trip
//     ends: what Bustrip::get_busv() holds after the run says nothing about the length the links saw).
#pragma once
void mz_hook_stop_visit(int vid, int trip_id, int stop_id, int link_id, double enter_time, double exit_time);
void mz_hook_bus_dispatch(int vid, int trip_id, double dispatch_time, double vehicle_length);
#define set_enter_time(t)                                                                                                   \
    set_enter_time((mz_hook_stop_visit(entering_trip->get_busv()->get_id(), entering_trip->get_id(), this->get_id(),       \
                                       this->get_link_id(), (t), exit_time),                                               \
                    (t)))
#define insert_veh(v, t) insert_veh((mz_hook_bus_dispatch((v)->get_id(), this->get_id(), (t), (double)(v)->get_length()), (v)), (t))

// TEST INFRASTRUCTURE: compile-time hook, force-included ONLY when compiling the reference's link.cpp in place
// (after vehicle.h, so the declaration of Vehicle::add_meters is untouched). Both successful-exit sites of the
// reference call veh->add_meters(length) (B/link.cpp:579 and 646) with `veh`, `time` and `this` in scope; the macro
// logs (vehicle id, link id, entry time, exit time) without changing a byte of the reference sources.
#pragma once
void mz_hook_exit(int vid, int lid, double entry_time, double exit_time);
#define add_meters(len) add_meters((mz_hook_exit(veh->get_id(), this->get_id(), veh->get_entry_time(), time), (len)))

// TEST INFRASTRUCTURE: force-included ONLY when compiling the reference's od.cpp in place (after node.h): logs every
// generated trip at the single call site Origin::insert_veh(veh,time) in ODaction::execute (B/od.cpp:84).
#pragma once
class Vehicle;
void mz_hook_trip(Vehicle *veh, double time);
#define insert_veh(v, t) insert_veh((mz_hook_trip((v), (t)), (v)), (t))

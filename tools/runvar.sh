for t in ns0 ns400; do cp mezzo_b200/libmezzo_b200.so /tmp/orig.so; cp tools/_var/lib_$t.so mezzo_b200/libmezzo_b200.so; echo "variant $t"; timeout -s KILL 200 python tools/sim_sweep.py sim_grid100 1000000 0 2>&1 | tail -1; cp /tmp/orig.so mezzo_b200/libmezzo_b200.so; done
echo baseline; timeout -s KILL 200 python tools/sim_sweep.py sim_grid100 1000000 0 2>&1 | tail -1

#!/bin/bash
# Run a python script with the in-tree library ("new") and with tmp_ab/old.so ("old").
set -u
cp stringwars_b200/lib/libstringwars_b200.so /tmp/new.so
echo "== new"; python "$@" 2>&1 | grep -v Warning | grep -v "pad = "
cp tmp_ab/old.so stringwars_b200/lib/libstringwars_b200.so
echo "== old"; python "$@" 2>&1 | grep -v Warning | grep -v "pad = "
cp /tmp/new.so stringwars_b200/lib/libstringwars_b200.so

bash profiles/run_r03c_ab.sh r03f "list4pfu4" 2>&1 | grep -v "^+"
bash profiles/run_prof_r03.sh r03f > /dev/null 2>&1

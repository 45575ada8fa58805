for sl in 0 40 200 1000; do
  echo "== CARC_S3F_SLEEP=$sl"
  CARC_S3F_SLEEP=$sl timeout 300 python scripts/matvec_paths.py --paths 3 --sizes 4:16,6:16,8:16 --out /tmp/o.md > /tmp/o.log 2>&1
  sort -u /tmp/o.md | grep -v "^| D\|^|--"
done

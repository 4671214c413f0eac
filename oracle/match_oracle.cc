// ORACLE — test infrastructure only.  Matcher restatement (filled in below).

# package init filled in later

bash tools/scout_ab.sh $1 && bash tools/prof_games.sh $1 scout
